#!/bin/bash
# Round 2, GPU call 11: stream BatchNorm kernels -- micro-benchmark A/B, train tests, train step.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/bench_bn.py --batch 16 > gpurun_out/r2k_bench_bn.jsonl 2> gpurun_out/r2k_bench_bn.err; grep ms_per_step gpurun_out/r2k_bench_bn.jsonl; tail -2 gpurun_out/r2k_bench_bn.err
timeout 900 python -m pytest tests/test_zz_gpu_train.py -x -q > gpurun_out/r2k_pytest_train.log 2>&1; tail -4 gpurun_out/r2k_pytest_train.log
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2k_train_b16.json 2> gpurun_out/r2k_train_b16.err
cut -c1-900 gpurun_out/r2k_train_b16.json; tail -2 gpurun_out/r2k_train_b16.err
