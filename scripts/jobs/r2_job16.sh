#!/bin/bash
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -q -x -k "forward_backward or golden or test_gpu_model or families or batched" > gpurun_out/j16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j16_pytest.log
timeout 200 python scripts/run_decode_once.py > gpurun_out/j16_decode.log 2>&1
CMD="python scripts/run_decode_once.py"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j16_launches_decode.csv $CMD > gpurun_out/j16_ncu_decode.log 2>&1
for k in bwd_bulk_kernel atb_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o /tmp/prof/dec_$k $CMD > gpurun_out/j16_ncufull_dec_$k.log 2>&1
done
CMD2="python bench.py --workload stress_dense --bins 200000 --steps 1 --warmup 2 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_step_gemm -s 3000 -c 1 -o /tmp/prof/dense_gemm $CMD2 > gpurun_out/j16_ncufull_dense_gemm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_bwd_update -s 2000 -c 1 -o /tmp/prof/dense_bwd_update $CMD2 > gpurun_out/j16_ncufull_dense_bwd.log 2>&1
python scripts/ncu_summary.py /tmp/prof/*.ncu-rep > gpurun_out/j16_ncu_full_summary.csv 2> gpurun_out/j16_ncu_summary.err
for f in /tmp/prof/dense_gemm.ncu-rep /tmp/prof/dec_bwd_bulk_kernel.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page source --csv --print-source cuda,sass > /tmp/prof/src_$b.csv 2>/dev/null
  python scripts/ncu_src_lines.py /tmp/prof/src_$b.csv > gpurun_out/j16_hot_lines_$b.txt 2>&1
done
tail -n 3 gpurun_out/j16_pytest.log; cat gpurun_out/j16_decode.log
