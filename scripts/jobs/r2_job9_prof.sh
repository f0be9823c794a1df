#!/bin/bash
mkdir -p gpurun_out /tmp/prof
CMD="python bench.py --steps 2 --warmup 8 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
$CMD > gpurun_out/j9_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/j9_launches_headline.csv $CMD > gpurun_out/j9_ncu1.log 2>&1
for k in fwd_c_kernel bwd_c_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 2 -o /tmp/prof/$k $CMD > gpurun_out/j9_ncufull_$k.log 2>&1
done
timeout 400 ncu --set full --clock-control none -k regex:mstep_adam_lag_kernel -s 6 -c 1 -o /tmp/prof/mstep $CMD > gpurun_out/j9_ncufull_mstep.log 2>&1
python scripts/ncu_summary.py /tmp/prof/*.ncu-rep > gpurun_out/j9_ncu_full_summary.csv 2> gpurun_out/j9_ncu_summary.err
for f in /tmp/prof/fwd_c_kernel.ncu-rep /tmp/prof/bwd_c_kernel.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page source --csv --print-source cuda,sass > /tmp/prof/src_$b.csv 2>/dev/null
  python scripts/ncu_src_lines.py /tmp/prof/src_$b.csv > gpurun_out/j9_hot_lines_$b.txt 2>&1
done
# the knobs of the pair emission kernel once more, now with the occupancy query cached (host time out of the phase)
timeout 200 python bench.py --steps 10 --warmup 5 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j9_bench_quick.json 2> gpurun_out/j9_bench_quick.err
du -sh gpurun_out > gpurun_out/j9_sizes.txt
