#!/bin/bash
# Round 2, GPU call 12: launch list of one eager B=16 train1 step under ncu (shares only).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2l_train_launches.csv \
  python tools/bench_train.py --batch 16 --mode flat --steps 1 --warmup 0 > gpurun_out/r2l_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2l_train_launches.csv "one eager B=16 train1 step under ncu (round 2, call 12; includes one-time weight packing)" > gpurun_out/r2l_train_launches_summary.md 2>&1 || true
head -60 gpurun_out/r2l_train_launches_summary.md
