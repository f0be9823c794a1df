#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
r() { env "$@" timeout 120 python scripts/bench_emission.py 2>&1 | tail -1; }
r TAG=v3

r TAG=v3_st1 PMG_EM_STAGGER=0
r TAG=v1 PMG_EM_KERNEL=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "emission or naive or atb or stat" > gpurun_out/pytest_h1.log 2>&1; echo "kernel tests rc=$?"
tail -5 gpurun_out/pytest_h1.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-decode > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_h.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
print('host', {k:round(v,3) for k,v in d['phases_host_ms_per_step'].items()})
print(d['clocks'])
PY
