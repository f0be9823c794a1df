#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-decode --no-e2e > gpurun_out/bench_r_8gpu.json 2> gpurun_out/bench_r_8gpu.err; echo "bench8 rc=$?"
tail -2 gpurun_out/bench_r_8gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'], d['config']['adam_steps_per_iter'])
PY
