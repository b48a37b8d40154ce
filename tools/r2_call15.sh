#!/bin/bash
# Round 2, GPU call 16: crop kernel v2 + single-arena staging (tests + bench), train step with the restored BN reduce grid.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_processer.py -m gpu -x -q > gpurun_out/r2p_pytest_processer.log 2>&1; tail -3 gpurun_out/r2p_pytest_processer.log
timeout 300 python tools/bench_crop.py --batch 64 --steps 10 > gpurun_out/r2p_bench_crop.json 2> gpurun_out/r2p_bench_crop.err; cat gpurun_out/r2p_bench_crop.json; tail -3 gpurun_out/r2p_bench_crop.err
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2p_train_b16.json 2> gpurun_out/r2p_train_b16.err
cut -c1-400 gpurun_out/r2p_train_b16.json; tail -2 gpurun_out/r2p_train_b16.err
