#!/bin/bash
# Round 2, GPU call 31: train3 step as a CUDA graph (device-scheduled RAdam) -- tests + bench in both modes.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_zz_gpu_train.py tests/test_optimizer.py -x -q > gpurun_out/r2af_pytest_train.log 2>&1; tail -6 gpurun_out/r2af_pytest_train.log
timeout 300 python tools/bench_train3.py --batch 64 --steps 5 --warmup 2 --mode graph > gpurun_out/r2af_train3_graph.json 2> gpurun_out/r2af_train3_graph.err; cut -c1-420 gpurun_out/r2af_train3_graph.json; tail -2 gpurun_out/r2af_train3_graph.err
timeout 300 python tools/bench_train3.py --batch 64 --steps 5 --warmup 2 --mode eager 2>/dev/null | cut -c1-300
