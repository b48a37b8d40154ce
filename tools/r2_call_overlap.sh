#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_page.py tests/test_gpu_backend_abi.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-train1 --no-side --no-gpu-reference --no-cpu-baseline > gpurun_out/r02al_bench_overlap.json 2> gpurun_out/r02al_bench_overlap.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02al_bench_overlap.json') if l.startswith('{')][-1])
print('value', round(d['value'],1), 'e2e', d['e2e'], d['clocks'])
PY
tail -n 3 gpurun_out/r02al_bench_overlap.err
