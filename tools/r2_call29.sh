#!/bin/bash
# Round 2, GPU call 30: fused attention backward -- train tests, train3 step (bench + warm kernel-time table).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_zz_gpu_train.py -x -q > gpurun_out/r2ae_pytest_train.log 2>&1; tail -3 gpurun_out/r2ae_pytest_train.log
timeout 300 python tools/bench_train3.py --batch 64 --steps 3 --warmup 2 > gpurun_out/r2ae_train3.json 2>/dev/null; cut -c1-400 gpurun_out/r2ae_train3.json
timeout 300 python tools/profile_train3_torch.py 64 > gpurun_out/r2ae_train3_kernels.md 2>/dev/null; head -24 gpurun_out/r2ae_train3_kernels.md | cut -c1-130
