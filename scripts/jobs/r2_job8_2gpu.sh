#!/bin/bash
# 2 x B200: NCCL path of the time-sharded fit (tests + strong/weak bench lines + graph-with-NCCL check)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j8_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/j8_pytest_multirank.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j8_pytest_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/j8_bench_2gpu_strong.json 2> gpurun_out/j8_bench_2gpu_strong.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 10 --scaling weak --no-e2e > gpurun_out/j8_bench_2gpu_weak.json 2> gpurun_out/j8_bench_2gpu_weak.err
PMG_EM_GRAPH=1 PMG_EM_GRAPH_DIST=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 25 --no-e2e --no-parity > gpurun_out/j8_bench_2gpu_graph.json 2> gpurun_out/j8_bench_2gpu_graph.err
timeout 300 $TR bench.py --gpus 2 --workload nb --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j8_bench_2gpu_nb.json 2> gpurun_out/j8_bench_2gpu_nb.err
tail -n 3 gpurun_out/j8_pytest_multirank.log
