#!/bin/bash
# round 2, GPU job 1: multirank test on one GPU (gloo), seam errors vs halo, 8 ranks emulated on one GPU, pooled fit
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j1_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/j1_pytest_multirank.log 2>&1
timeout 300 python scripts/diag_seams.py halo 1000000 > gpurun_out/j1_halo_1e6.log 2>&1
timeout 300 python scripts/diag_seams.py halo 125000 > gpurun_out/j1_halo_125k.log 2>&1
BACKEND=gloo ITERS=14 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/diag_seams.py ranks 1000000 > gpurun_out/j1_ranks8_gloo.log 2>&1
ITERS=14 timeout 400 python scripts/diag_seams.py pooled 1000000 8 > gpurun_out/j1_pooled8.log 2>&1
tail -3 gpurun_out/j1_*.log
