#!/bin/bash
mkdir -p gpurun_out
P=29617
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/j17_pytest_multirank.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j17_pytest_multirank.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/j17_bench_8gpu_strong.json 2> gpurun_out/j17_bench_8gpu_strong.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-decode --no-cpu-baseline ) > gpurun_out/j17_bench_4gpu_strong.json 2> gpurun_out/j17_bench_4gpu_strong.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+3)) bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-decode --no-cpu-baseline ) > gpurun_out/j17_bench_2gpu_strong.json 2> gpurun_out/j17_bench_2gpu_strong.err
tail -n 3 gpurun_out/j17_pytest_multirank.log
