#!/bin/bash
# Round 2, GPU call 2: suite after the determinism / graph / SE-kernel changes, then the train step as a CUDA graph.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -q -m gpu --tb=short -rxX -p no:cacheprovider -x > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2b_pytest_gpu.log
for cfg in "2 graph" "2 flat" "16 graph" "16 flat"; do
  set -- $cfg
  timeout 400 python tools/bench_train.py --batch $1 --mode $2 --steps 3 --warmup 1 > gpurun_out/r2b_train_b$1_$2.json 2> gpurun_out/r2b_train_b$1_$2.err
  cat gpurun_out/r2b_train_b$1_$2.json; tail -2 gpurun_out/r2b_train_b$1_$2.err
done
# GPU-time shares of one eager B=8 step
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2b_train_launches.csv \
  python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2b_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2b_train_launches.csv "one eager B=8 train1 step (mma.sync wgrad) under ncu" > gpurun_out/r2b_train_launches_summary.md 2>&1 || true
head -45 gpurun_out/r2b_train_launches_summary.md
