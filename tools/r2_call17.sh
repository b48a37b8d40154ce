#!/bin/bash
# Round 2, GPU call 17: distortion kernels v2 (tests + bench), ncu --set full of the attention kernel (cfg4) and of the crop kernels.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_processer.py -m gpu -x -q > gpurun_out/r2q_pytest_processer.log 2>&1; tail -2 gpurun_out/r2q_pytest_processer.log
timeout 300 python tools/bench_crop.py --batch 64 --steps 10 > gpurun_out/r2q_bench_crop.json 2> gpurun_out/r2q_bench_crop.err; cat gpurun_out/r2q_bench_crop.json; tail -2 gpurun_out/r2q_bench_crop.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_mma_kernel --launch-skip 60 -c 2 -o gpurun_out/r2q_ncu_attention -f \
  python tools/bench_transformer.py cfg4 bf16 > gpurun_out/r2q_ncu_attention.log 2>&1; tail -2 gpurun_out/r2q_ncu_attention.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_ --launch-skip 10 -c 5 -o gpurun_out/r2q_ncu_crop -f \
  python tools/bench_crop.py --batch 64 --steps 3 --no-cpu > gpurun_out/r2q_ncu_crop.log 2>&1; tail -2 gpurun_out/r2q_ncu_crop.log
ls -la gpurun_out/*.ncu-rep | tail -3
