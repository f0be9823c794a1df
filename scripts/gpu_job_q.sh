#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_q.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
print('e2e', d['e2e']['value'], d['e2e']['wall_s']); print('decode', d['decode']); print(d['clocks']); print(d['cpu_baseline']); print(d['roofline'])
PY
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-decode --phase-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_q.csv $B > gpurun_out/ncu_launch_q.log 2>&1; echo "ncu launches rc=$?"
B="python bench.py --steps 2 --warmup 40 --no-e2e --no-cpu-baseline --no-decode --phase-steps 0"
for k in emission_tc_kernel atb_tc_kernel fwd_c_kernel bwd_c_kernel mstep_adam_lag_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 36 -c 1 \
     -o gpurun_out/prof_q_$k -f $B > gpurun_out/ncu_full_q_$k.log 2>&1; echo "ncu $k rc=$? $(grep -c Profiling gpurun_out/ncu_full_q_$k.log)"
done
