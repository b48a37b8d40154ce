#!/bin/bash
# Round 2, GPU call 9: validation of HEAD -- smoke, the whole GPU suite, the full bench.py line.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1; tail -2 gpurun_out/r2i_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
cut -c1-1200 gpurun_out/r2i_bench_n1.json; tail -3 gpurun_out/r2i_bench_n1.err
