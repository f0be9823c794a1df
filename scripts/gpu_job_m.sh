#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/pytest_m.log 2>&1; echo "multirank tests rc=$?"
tail -3 gpurun_out/pytest_m.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m_2gpu.json 2> gpurun_out/bench_m_2gpu.err; echo "bench2 rc=$?"
tail -2 gpurun_out/bench_m_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_m_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['e2e'] and (d['e2e']['value'], d['e2e']['wall_s']))
PY
