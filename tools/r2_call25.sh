#!/bin/bash
# Round 2, GPU call 25: re-validation after the dead-path removal (whole GPU suite + short bench).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2x_smoke.log 2>&1; tail -1 gpurun_out/r2x_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2x_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; cut -c1-300 gpurun_out/r2x_bench_n1.json; tail -2 gpurun_out/r2x_bench_n1.err
