#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/j14_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j14_pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/j14_bench_default.json 2> gpurun_out/j14_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/j14_bench_ref.json 2> gpurun_out/j14_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j14_smoke.log 2>&1
tail -n 4 gpurun_out/j14_pytest_gpu.log; tail -n 2 gpurun_out/j14_smoke.log
