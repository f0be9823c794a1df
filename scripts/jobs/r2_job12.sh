#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_dense_scan.py > gpurun_out/j12_diag_dense.log 2>&1
for ms in 20 200 1000; do
  timeout 200 python bench.py --workload nb --bins 1250000 --steps 20 --warmup 3 --no-cpu-baseline --clocks-ms $ms > gpurun_out/j12_nb_clk$ms.json 2> gpurun_out/j12_nb_clk$ms.err
done
for ms in 20 1000; do
  timeout 300 python bench.py --steps 20 --warmup 8 --no-e2e --no-decode --no-cpu-baseline --phase-steps 0 --clocks-ms $ms > gpurun_out/j12_em_clk$ms.json 2> gpurun_out/j12_em_clk$ms.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "session_shape_default" > gpurun_out/j12_pytest.log 2>&1
CMD="python scripts/run_decode_once.py"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j12_launches_decode.csv $CMD > gpurun_out/j12_ncu_decode.log 2>&1
cat gpurun_out/j12_diag_dense.log; tail -3 gpurun_out/j12_pytest.log
