#!/bin/bash
# Round 2, GPU call 23: batched loads in dw_wgrad / spatial_sum / scale_bc / se_fc_bwd1, fewer reduce chunks: tests + train step.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_zz_gpu_train.py -x -q > gpurun_out/r2u_pytest_train.log 2>&1; tail -4 gpurun_out/r2u_pytest_train.log
timeout 400 python tools/bench_train.py --batch 16 --mode graph --steps 3 --warmup 1 > gpurun_out/r2u_train_b16.json 2> gpurun_out/r2u_train_b16.err
cut -c1-400 gpurun_out/r2u_train_b16.json; tail -2 gpurun_out/r2u_train_b16.err
timeout 600 python tools/profile_train_torch.py --batch 16 --steps 1 > gpurun_out/r2u_train_kernels.md 2> gpurun_out/r2u_train_kernels.err
grep -n "dw_wgrad\|se_fc_bwd1\|spatial_sum\|scale_bc\|step wall" gpurun_out/r2u_train_kernels.md | cut -c1-140
