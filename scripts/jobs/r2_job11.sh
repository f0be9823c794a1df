#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lockstep or stress_shape or session_shape_default" > gpurun_out/j11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j11_pytest.log
T=20000 timeout 600 python scripts/bench_dense_scan.py > gpurun_out/j11_dense_scan.log 2>&1
timeout 900 python bench.py --workload stress_dense --steps 3 --warmup 3 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j11_bench_dense.json 2> gpurun_out/j11_bench_dense.err
timeout 300 python scripts/profile_nb_host.py > gpurun_out/j11_nb_host.log 2>&1
tail -n 5 gpurun_out/j11_pytest.log; cat gpurun_out/j11_dense_scan.log
