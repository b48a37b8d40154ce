#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): page pipeline tests, graph counter debug, bench.py at N=2 with the NCCL train step, N=1 bench.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_page.py tests/test_zzzz_gpu_wgrad_tc.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/r2d_page_tests.log 2>&1
tail -25 gpurun_out/r2d_page_tests.log
timeout 200 python tools/debug_graph_counters.py > gpurun_out/r2d_counters.log 2>&1; tail -8 gpurun_out/r2d_counters.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
cat gpurun_out/r2d_bench_n2.json; tail -5 gpurun_out/r2d_bench_n2.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
cat gpurun_out/r2d_bench_n1.json; tail -5 gpurun_out/r2d_bench_n1.err
