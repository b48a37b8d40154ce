#!/bin/bash
# Round 2, GPU call 6: new tests (predict_each, train kernels after the BN unroll), train step, page bench, BN kernel profile.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_backend_abi.py tests/test_zz_gpu_train.py tests/test_zzz_gpu_wgrad_mma.py tests/test_gpu_transformer.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
tail -15 gpurun_out/r2f_pytest.log
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2f_train_b16.json 2> gpurun_out/r2f_train_b16.err
cat gpurun_out/r2f_train_b16.json; tail -2 gpurun_out/r2f_train_b16.err
timeout 300 python tools/bench_page.py --pages 3 --chunks 32 > gpurun_out/r2f_page.json 2> gpurun_out/r2f_page.err; cat gpurun_out/r2f_page.json; tail -3 gpurun_out/r2f_page.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2f_train_launches.csv \
  python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2f_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2f_train_launches.csv "one eager B=8 train1 step under ncu (round 2, call 6)" > gpurun_out/r2f_train_launches_summary.md 2>&1 || true
head -24 gpurun_out/r2f_train_launches_summary.md
