#!/bin/bash
mkdir -p gpurun_out
PMG_EM_PAIR=0 timeout 200 python scripts/check_emission_pair.py run single > gpurun_out/j4_empair.log 2>&1
PMG_EM_PAIR=1 timeout 200 python scripts/check_emission_pair.py run pair >> gpurun_out/j4_empair.log 2>&1
echo "pair rc=$?" >> gpurun_out/j4_empair.log
python scripts/check_emission_pair.py cmp single pair >> gpurun_out/j4_empair.log 2>&1
rm -f gpurun_out/ll_*.npy
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j4_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j4_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err
timeout 300 python bench.py --steps 20 --warmup 10 --bins 125000 --no-e2e --no-decode --no-cpu-baseline > gpurun_out/j4_bench_125k.json 2> gpurun_out/j4_bench_125k.err
PMG_BENCH_BACKEND=gloo timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 10 --no-e2e > gpurun_out/j4_bench_gloo8.json 2> gpurun_out/j4_bench_gloo8.err
timeout 300 python bench.py --workload nb --bins 1250000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j4_bench_nb.json 2> gpurun_out/j4_bench_nb.err
tail -n 3 gpurun_out/j4_pytest_gpu.log; cat gpurun_out/j4_empair.log
