#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_families.py tests/test_gpu_batched.py -m gpu -q > gpurun_out/j6_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j6_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j6_bench.json 2> gpurun_out/j6_bench.err
timeout 300 python bench.py --steps 20 --warmup 10 --bins 125000 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j6_bench_125k.json 2> gpurun_out/j6_bench_125k.err
PMG_EM_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 10 --bins 125000 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j6_bench_125k_nograph.json 2> gpurun_out/j6_bench_125k_nograph.err
PMG_EM_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j6_bench_nograph.json 2> gpurun_out/j6_bench_nograph.err
timeout 300 python bench.py --workload session --steps 30 --warmup 10 --no-decode --no-cpu-baseline > gpurun_out/j6_bench_session.json 2> gpurun_out/j6_bench_session.err
CMD="env PMG_EM_GRAPH=0 python bench.py --steps 2 --warmup 8 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
$CMD > gpurun_out/j6_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 160 --csv --log-file gpurun_out/j6_launches_headline.csv $CMD > gpurun_out/j6_ncu1.log 2>&1
CMD2="env PMG_EM_GRAPH=0 python bench.py --steps 2 --warmup 12 --bins 125000 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
$CMD2 > gpurun_out/j6_plain2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 160 --csv --log-file gpurun_out/j6_launches_125k.csv $CMD2 > gpurun_out/j6_ncu2.log 2>&1
mkdir -p /tmp/prof
for k in emission_tc2_kernel fwd_c_kernel bwd_c_kernel atb_tc_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -o /tmp/prof/$k $CMD > gpurun_out/j6_ncufull_$k.log 2>&1
done
CMD3="python scripts/run_decode_once.py"
$CMD3 > gpurun_out/j6_decode_plain.log 2>&1 && for k in atb_tc_kernel bwd_bulk_kernel fwd_bulk_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o /tmp/prof/dec_$k $CMD3 > gpurun_out/j6_ncufulldec_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/j6_launches_decode.csv $CMD3 > gpurun_out/j6_ncu3.log 2>&1
# summaries here (the reports themselves are too large to bring back: 64 MiB limit)
python scripts/ncu_summary.py /tmp/prof/*.ncu-rep > gpurun_out/j6_ncu_full_summary.csv 2> gpurun_out/j6_ncu_summary.err
for f in /tmp/prof/*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page details --csv > gpurun_out/j6_details_$b.csv 2>/dev/null
  ncu -i $f --page source --csv > /tmp/prof/src_$b.csv 2>/dev/null
  python scripts/ncu_src_top.py /tmp/prof/src_$b.csv > gpurun_out/j6_hot_$b.txt 2>&1
done
ls -la /tmp/prof > gpurun_out/j6_prof_sizes.txt
cp /tmp/prof/emission_tc2_kernel.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out >> gpurun_out/j6_prof_sizes.txt
tail -n 3 gpurun_out/j6_pytest_gpu.log
