#!/bin/bash
# Round 2, GPU call 8: train step after the BN grid fix, BN micro-benchmark again, full bench.py line.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 400 python tools/bench_bn.py --batch 16 > gpurun_out/r2h_bench_bn.jsonl 2> gpurun_out/r2h_bench_bn.err; head -1 gpurun_out/r2h_bench_bn.jsonl; tail -2 gpurun_out/r2h_bench_bn.err
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2h_train_b16.json 2> gpurun_out/r2h_train_b16.err
cat gpurun_out/r2h_train_b16.json; tail -2 gpurun_out/r2h_train_b16.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
cat gpurun_out/r2h_bench_n1.json | cut -c1-1500; tail -3 gpurun_out/r2h_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2h_train_launches.csv \
  python tools/bench_train.py --batch 8 --mode flat --steps 2 --warmup 0 > gpurun_out/r2h_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2h_train_launches.csv "two eager B=8 train1 steps under ncu (round 2, call 8)" > gpurun_out/r2h_train_launches_summary.md 2>&1 || true
head -30 gpurun_out/r2h_train_launches_summary.md
