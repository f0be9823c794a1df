#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j3_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j3_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/j3_bench.json 2> gpurun_out/j3_bench.err
BACKEND=gloo ITERS=14 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/diag_seams.py ranks 1000000 > gpurun_out/j3_ranks8_gloo.log 2>&1
ITERS=20 timeout 300 python scripts/diag_seams.py pooled 125000 1 > gpurun_out/j3_single_125k.log 2>&1
tail -n 3 gpurun_out/j3_pytest_gpu.log
