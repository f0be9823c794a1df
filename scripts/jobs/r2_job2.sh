#!/bin/bash
mkdir -p gpurun_out
DISSECT_FROM=9 ITERS=12 timeout 400 python scripts/diag_seams.py pooled 1000000 8 > gpurun_out/j2_pooled8_dissect.log 2>&1
BACKEND=gloo DISSECT_FROM=9 ITERS=11 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/diag_seams.py ranks 1000000 > gpurun_out/j2_ranks8_dissect.log 2>&1
echo done
