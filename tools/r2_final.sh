#!/bin/bash
# Round 2 final evidence on one B200: smoke, whole GPU suite, bench line (ours + reference arm), ncu launch lists (bench.py itself and
# two forwards with DRAM bytes per launch), ncu --set full of the dominant kernels.  Outputs under gpurun_out/r2z_*.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; cut -c1-500 gpurun_out/r2z_bench_n1.json; tail -2 gpurun_out/r2z_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; cut -c1-300 gpurun_out/r2z_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r2z_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-train1 --no-side --no-gpu-reference --no-cpu-baseline > gpurun_out/r2z_bench_under_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2z_bench_launches.csv "bench.py --steps 2 --warmup 3 (detector forward B=32 bf16 + e2e) under ncu" > gpurun_out/r2z_bench_launches_summary.md 2>&1; head -24 gpurun_out/r2z_bench_launches_summary.md
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2z_launches.csv python tools/run_forward.py 32 2 > gpurun_out/r2z_launch_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_launches.csv gpurun_out/r2z_launches_summary.md gpurun_out/r2z_traffic.json > /dev/null 2>&1; cat gpurun_out/r2z_traffic.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tma -s 197 -c 2 -o gpurun_out/r2z_ncu_heads -f python tools/run_forward.py 32 1 > gpurun_out/r2z_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tma -s 70 -c 2 -o gpurun_out/r2z_ncu_mbconv -f python tools/run_forward.py 32 1 > gpurun_out/r2z_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_stream -s 40 -c 1 -o gpurun_out/r2z_ncu_bn_bwd -f python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2z_ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc_kernel -s 40 -c 1 -o gpurun_out/r2z_ncu_wgrad -f python tools/bench_train.py --batch 8 --mode flat --steps 1 --warmup 0 > gpurun_out/r2z_ncu_d.log 2>&1
ls -la gpurun_out/r2z_*.ncu-rep
