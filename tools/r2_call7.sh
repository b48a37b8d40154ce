#!/bin/bash
# Round 2, GPU call 7: train tests, BN micro-benchmark per unroll factor, ncu --set full of the top train-step kernels.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_zz_gpu_train.py -q -m gpu --tb=short -p no:cacheprovider -k "graph or checkpoint or bn or train_step" -s > gpurun_out/r2g_pytest.log 2>&1
tail -8 gpurun_out/r2g_pytest.log
timeout 400 python tools/bench_bn.py --batch 16 > gpurun_out/r2g_bench_bn.jsonl 2> gpurun_out/r2g_bench_bn.err; head -3 gpurun_out/r2g_bench_bn.jsonl; tail -2 gpurun_out/r2g_bench_bn.err
for k in bn_act_bwd_vec_kernel col_reduce_vec_kernel bn_act_vec_kernel conv_wgrad_tc_kernel dw_wgrad_vec_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 60 --launch-count 3 -f -o gpurun_out/r2g_ncu_$k \
    python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2g_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2g_ncu_$k.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_$k.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
