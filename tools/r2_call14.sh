#!/bin/bash
# Round 2, GPU call 14: warm kernel-time table of the eager B=16 train step (torch.profiler).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python tools/profile_train_torch.py --batch 16 --steps 1 > gpurun_out/r2n_train_kernels.md 2> gpurun_out/r2n_train_kernels.err
head -56 gpurun_out/r2n_train_kernels.md; tail -3 gpurun_out/r2n_train_kernels.err
