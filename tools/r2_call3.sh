#!/bin/bash
# Round 2, GPU call 3: tcgen05 weight gradient parity (both modes, own process first), suite, train step with it.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_zzzz_gpu_wgrad_tc.py -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r2c_wgrad_tc.log 2>&1
tail -30 gpurun_out/r2c_wgrad_tc.log
timeout 900 python -m pytest tests -q -m gpu --tb=short -rxX -p no:cacheprovider --deselect tests/test_zzzz_gpu_wgrad_tc.py > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2c_pytest_gpu.log
for m in 0 1 2; do
  FTC_WGRAD_TC=$m timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2c_train_b16_tc$m.json 2> gpurun_out/r2c_train_b16_tc$m.err
  cat gpurun_out/r2c_train_b16_tc$m.json; tail -2 gpurun_out/r2c_train_b16_tc$m.err
done
FTC_WGRAD_TC=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2c_train_launches.csv \
  python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2c_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2c_train_launches.csv "one eager B=8 train1 step (tcgen05 wgrad mode 1) under ncu" > gpurun_out/r2c_train_launches_summary.md 2>&1 || true
head -40 gpurun_out/r2c_train_launches_summary.md
