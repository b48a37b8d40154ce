#!/bin/bash
# Round 2, GPU call 5: full suite (green check), train step after the depthwise strip / wgrad / running-stat changes, profile.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -q -m gpu --tb=short -rxX -p no:cacheprovider > gpurun_out/r2e_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2e_pytest_gpu.log
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2e_train_b16.json 2> gpurun_out/r2e_train_b16.err
cat gpurun_out/r2e_train_b16.json; tail -2 gpurun_out/r2e_train_b16.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2e_train_launches.csv \
  python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2e_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2e_train_launches.csv "one eager B=8 train1 step under ncu (round 2, call 5)" > gpurun_out/r2e_train_launches_summary.md 2>&1 || true
head -40 gpurun_out/r2e_train_launches_summary.md
