#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/pytest_p.log 2>&1; echo "multirank tests rc=$?"
tail -3 gpurun_out/pytest_p.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench_p_4gpu.json 2> gpurun_out/bench_p_4gpu.err; echo "bench4 rc=$?"
tail -2 gpurun_out/bench_p_4gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_p_4gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'], d['config']['adam_steps_per_iter'], d['e2e'] and (d['e2e']['value'], d['e2e']['wall_s']))
PY
