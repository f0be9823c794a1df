#!/bin/bash
# experiments: emission pipeline depth / tile width, scan halo; golden GPU tests; ncu --set full of the two GEMMs
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_b.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_b_$name.json 2> gpurun_out/bench_b_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_b_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'])
except Exception as e: print('$name failed', e)
PY
}
run default A=1
run em_st2 PMG_EM_STAGES=2
run em_nt4 PMG_EM_NT=4
run em_nt3 PMG_EM_NT=3
run halo128 PMG_HALO=128
run halo64 PMG_HALO=64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"emission_tc_kernel|atb_tc_kernel" --launch-skip 2 -c 2 \
   -o gpurun_out/prof_b_gemm -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
