#!/bin/bash
# Round 2, GPU call 10: f3 input-pipeline kernels on hardware (parity + bench), depthwise spill fix check, page split.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_processer.py -m gpu -x -q > gpurun_out/r2j_pytest_processer.log 2>&1; tail -5 gpurun_out/r2j_pytest_processer.log
timeout 300 python tools/bench_crop.py --batch 64 --steps 10 > gpurun_out/r2j_bench_crop.json 2> gpurun_out/r2j_bench_crop.err; cat gpurun_out/r2j_bench_crop.json; tail -3 gpurun_out/r2j_bench_crop.err
timeout 300 python tools/bench_page.py --pages 3 --chunks 32 --split > gpurun_out/r2j_page.json 2> gpurun_out/r2j_page.err; cat gpurun_out/r2j_page.json; tail -3 gpurun_out/r2j_page.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-side --no-gpu-reference --no-cpu-baseline > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
cut -c1-600 gpurun_out/r2j_bench_n1.json; tail -3 gpurun_out/r2j_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['per_kind_ms'])
print(d.get('train1',{}).get('value'), d.get('train1',{}).get('ms_per_step'))
PY
