#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "2 2" "3 2" "2 3"; do
  set -- $cfg
  PMG_EM2_NBUF=$1 PMG_EM2_STAGES=$2 TAG="nbuf$1_st$2" timeout 120 python scripts/bench_emission.py >> gpurun_out/j7_emission_knobs.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j7_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j7_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j7_bench.json 2> gpurun_out/j7_bench.err
timeout 300 python bench.py --steps 20 --warmup 15 --bins 125000 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j7_bench_125k.json 2> gpurun_out/j7_bench_125k.err
PMG_EM_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 25 --bins 125000 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j7_bench_125k_graph.json 2> gpurun_out/j7_bench_125k_graph.err
timeout 300 python bench.py --workload session --steps 30 --warmup 15 --no-decode --no-cpu-baseline > gpurun_out/j7_bench_session.json 2> gpurun_out/j7_bench_session.err
timeout 300 python bench.py --workload nb --bins 1250000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j7_bench_nb.json 2> gpurun_out/j7_bench_nb.err
timeout 200 python scripts/run_decode_once.py > gpurun_out/j7_decode.log 2>&1
CMD2="python bench.py --steps 2 --warmup 16 --bins 125000 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
$CMD2 > gpurun_out/j7_plain2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 120 --csv --log-file gpurun_out/j7_launches_125k.csv $CMD2 > gpurun_out/j7_ncu2.log 2>&1
tail -n 3 gpurun_out/j7_pytest_gpu.log; cat gpurun_out/j7_emission_knobs.log
