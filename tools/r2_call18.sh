#!/bin/bash
# Round 2, GPU call 18: whole GPU suite on HEAD (incl. the new full-size property tests), crop bench, train step with the fused dgrad pack.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2r_pytest_gpu.log
timeout 300 python tools/bench_crop.py --batch 64 --steps 10 > gpurun_out/r2r_bench_crop.json 2> gpurun_out/r2r_bench_crop.err; cut -c1-1400 gpurun_out/r2r_bench_crop.json; tail -2 gpurun_out/r2r_bench_crop.err
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2r_train_b16.json 2> gpurun_out/r2r_train_b16.err
cut -c1-400 gpurun_out/r2r_train_b16.json; tail -2 gpurun_out/r2r_train_b16.err
