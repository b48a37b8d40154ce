#!/bin/bash
# Ablation sweep of the per-op profile (tools/profile_ops.py) under the kernels' experiment env flags -> gpurun_out/ablate_*.txt
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/profile_ops.py 32 0.1 > gpurun_out/ablate_$name.txt 2>&1; echo "== $name: $(head -1 gpurun_out/ablate_$name.txt)"; sed -n 2,9p gpurun_out/ablate_$name.txt; }
run base A=0
run nostore FTC_TMA_FLAGS=32
run nose FTC_TMA_FLAGS=64
run nores FTC_TMA_FLAGS=128
run noepi FTC_TMA_FLAGS=256
run mt1 FTC_TMA_MT=1
run mt2 FTC_TMA_MT=2
run topsgemm FTC_TOPS_GEMM=1
FTC_TOPS_GEMM=1 timeout 600 python -m pytest tests/test_gpu_detector.py -m gpu -x -q 2>&1 | tail -3
