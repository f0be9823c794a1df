#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/pytest_f1.log 2>&1; echo "kernel tests rc=$?"
tail -8 gpurun_out/pytest_f1.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_f_$name.json 2> gpurun_out/bench_f_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_f_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'], d['config']['n_chain'])
except Exception as e: print('$name failed', e)
PY
}
run nw12 A=1
run nw8 PMG_EM_CHAINS_PER_SM=8
run nw12_st0 PMG_EM_STAGGER=0
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_f.log
PMG_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"
grep "fit_em timing" gpurun_out/bench_f.json; tail -3 gpurun_out/bench_f.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_c_kernel|bwd_c_kernel|emission_tc|atb_tc" --launch-skip 8 -c 4 \
   -o gpurun_out/prof_f -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_f.log 2>&1; echo "ncu full rc=$?"
grep Profiling gpurun_out/ncu_full_f.log | cut -c1-60
