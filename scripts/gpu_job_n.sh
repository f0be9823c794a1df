#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['phases_ms_per_step'].items()})
print('e2e', d['e2e']['value'], d['e2e']['wall_s']); print('decode', d['decode']); print(d['clocks']); print(d['roofline'])
PY
