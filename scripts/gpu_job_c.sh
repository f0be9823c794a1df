#!/bin/bash
# compact scan kernels: parity tests, full gpu suite, bench, ncu (source-level) of the two compact kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "compact" > gpurun_out/pytest_c1.log 2>&1; echo "compact tests rc=$?"
tail -15 gpurun_out/pytest_c1.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_c.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_c_$name.json 2> gpurun_out/bench_c_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'])
except Exception as e: print('$name failed', e)
PY
}
run compact A=1
run full PMG_SCAN_COMPACT=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_c_kernel|bwd_c_kernel" --launch-skip 2 -c 2 \
   -o gpurun_out/prof_c_scan -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
