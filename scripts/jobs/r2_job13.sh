#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_dense_scan.py > gpurun_out/j13_diag_dense.log 2>&1
timeout 300 python scripts/diag_atb_bias.py > gpurun_out/j13_diag_atb.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "lockstep or stress_shape" > gpurun_out/j13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j13_pytest.log
T=20000 timeout 600 python scripts/bench_dense_scan.py > gpurun_out/j13_dense_scan.log 2>&1
timeout 900 python bench.py --workload stress_dense --steps 3 --warmup 3 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j13_bench_dense.json 2> gpurun_out/j13_bench_dense.err
CMD="python bench.py --workload stress_dense --bins 200000 --steps 1 --warmup 2 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/j13_launches_dense.csv $CMD > gpurun_out/j13_ncu_dense.log 2>&1
cat gpurun_out/j13_diag_dense.log gpurun_out/j13_diag_atb.log; tail -3 gpurun_out/j13_pytest.log; cat gpurun_out/j13_dense_scan.log
