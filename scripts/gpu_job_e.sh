#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_e_$name.json 2> gpurun_out/bench_e_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_e_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()}, d['config']['seam_relays_in_timed_region'])
except Exception as e: print('$name failed', e)
PY
}
run stag1 PMG_EM_STAGGER=1
run stag0 PMG_EM_STAGGER=0
run stag0_st3 PMG_EM_STAGGER=0 PMG_EM_STAGES=3
run stag1_st3 PMG_EM_STAGGER=1 PMG_EM_STAGES=3
run full_stag0 PMG_SCAN_COMPACT=0 PMG_EM_STAGGER=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_e.log 2>&1; echo "ncu list rc=$?"
PMG_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"
grep "fit_em timing" gpurun_out/bench_e.json
tail -3 gpurun_out/bench_e.err
