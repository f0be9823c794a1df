#!/bin/bash
# r01 session-2 baseline job: gpu tests, bench with fit_em timing marks, ncu launch list, ncu --set full of the 4 hot kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc > gpurun_out/host_a.txt; free -g >> gpurun_out/host_a.txt; nvidia-smi -L >> gpurun_out/host_a.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"
PMG_TIMING=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
grep "fit_em timing" gpurun_out/bench_a.json gpurun_out/bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_a.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_a.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd_bulk|bwd_bulk|emission_tc|atb_tc" --launch-skip 4 -c 4 \
   -o gpurun_out/prof_a -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_a.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
