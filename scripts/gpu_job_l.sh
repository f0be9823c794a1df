#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 40 --no-e2e --no-cpu-baseline --no-decode --phase-steps 0"
for k in emission_tc_kernel atb_tc_kernel fwd_c_kernel bwd_c_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 36 -c 1 \
     -o gpurun_out/prof_l_$k -f $B > gpurun_out/ncu_full_l_$k.log 2>&1; echo "ncu $k rc=$?"
  grep -c Profiling gpurun_out/ncu_full_l_$k.log
done
