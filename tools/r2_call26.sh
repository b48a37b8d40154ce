#!/bin/bash
# Round 2, GPU call 26: depthwise strip kernel with several images per CTA -- detector / ops / train tests, A/B against one image per CTA.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2y_pytest_det.log 2>&1; tail -3 gpurun_out/r2y_pytest_det.log
for ipc in 1 0; do
  FTC_DW_IPC=$ipc timeout 400 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-side --no-train1 --no-cpu-baseline > gpurun_out/r2y_bench_ipc$ipc.json 2> gpurun_out/r2y_bench_ipc$ipc.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_ipc$ipc.json') if l.startswith('{')][-1])
print('ipc $ipc', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['per_kind_ms']['depthwise_se'], d['clocks']['sm_mhz'])
PY
done
timeout 900 python -m pytest tests/test_zz_gpu_train.py tests/test_gpu_page.py tests/test_gpu_backend_abi.py -x -q > gpurun_out/r2y_pytest_train.log 2>&1; tail -3 gpurun_out/r2y_pytest_train.log
