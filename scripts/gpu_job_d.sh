#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "compact" > gpurun_out/pytest_d1.log 2>&1; echo "compact tests rc=$?"
tail -15 gpurun_out/pytest_d1.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_d.log
PMG_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_d.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'], d['e2e'], d['config']['n_chain'])
    else: print(l)
PY
tail -3 gpurun_out/bench_d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_c_kernel|bwd_c_kernel" --launch-skip 4 -c 2 \
   -o gpurun_out/prof_d_scan -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_d.log 2>&1; echo "ncu full rc=$?"
grep -c Profiling gpurun_out/ncu_full_d.log
