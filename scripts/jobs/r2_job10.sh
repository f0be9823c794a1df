#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/j10_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j10_pytest_gpu.log
CMD="python bench.py --workload nb --bins 1250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/j10_nb_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j10_launches_nb.csv $CMD > gpurun_out/j10_ncu_nb.log 2>&1
timeout 300 python bench.py --workload stress --steps 3 --warmup 2 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j10_bench_stress.json 2> gpurun_out/j10_bench_stress.err
tail -n 3 gpurun_out/j10_pytest_gpu.log
