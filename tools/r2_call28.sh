#!/bin/bash
# Round 2, GPU call 28: SE excitation kernel on 512 threads -- detector / ops / page / ABI tests and the forward bench.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ops.py tests/test_gpu_page.py tests/test_gpu_backend_abi.py -m gpu -x -q > gpurun_out/r2ab_pytest_det.log 2>&1; tail -3 gpurun_out/r2ab_pytest_det.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-side --no-train1 --no-cpu-baseline > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2ab_bench.json') if l.startswith('{')][-1])
print(round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['per_kind_ms']['se_fc'], d['roofline']['per_kind_ms']['depthwise_se'], d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1))
PY
