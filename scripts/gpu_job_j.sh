#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
r() { env "$@" timeout 120 python scripts/bench_emission.py 2>&1 | tail -1; }
r TAG=v4

timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "emission or naive" > gpurun_out/pytest_j1.log 2>&1; echo "kernel tests rc=$?"
tail -3 gpurun_out/pytest_j1.log
PMG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_j.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
print('e2e', d['e2e']['value'], d['e2e']['wall_s'])
print('decode', d['decode'])
print(d['clocks'])
PY
grep "fit_em timing" gpurun_out/bench_j.json
