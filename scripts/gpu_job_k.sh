#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_k.log
timeout 900 python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_k.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_k.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
print('e2e', d['e2e']['value'], d['e2e']['wall_s']); print('decode', d['decode']); print(d['clocks']); print(d['cpu_baseline'])
PY
timeout 300 python scripts/diag_decode.py 2>&1 | tail -3
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-decode --phase-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_k.csv $B > gpurun_out/ncu_launch_k.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_c_kernel|bwd_c_kernel|emission_tc_kernel|atb_tc_kernel" --launch-skip 8 -c 4 \
   -o gpurun_out/prof_k -f $B > gpurun_out/ncu_full_k.log 2>&1; echo "ncu full rc=$?"
grep Profiling gpurun_out/ncu_full_k.log | cut -c1-70
