#!/bin/bash
# Round-end evidence: GPU tests, bench line, ncu launch list (time + DRAM bytes) and ncu --set full captures of the dominant
# kernels.  Outputs under gpurun_out/; tools/launch_summary.py and tools/ncu_summary.py turn them into profiles/*.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_tail.txt; cat gpurun_out/pytest_gpu_tail.txt
timeout 300 python tools/profile_ops.py 32 0.1 > gpurun_out/ops_profile.txt 2>&1; head -9 gpurun_out/ops_profile.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/run_forward.py 32 2 > gpurun_out/launch_list.log 2>&1
# level-3 head conv (the largest single kernel) + the feature top conv, and an MBConv expand / project pair
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tma -s 197 -c 2 -o gpurun_out/ncu_heads python tools/run_forward.py 32 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tma -s 70 -c 2 -o gpurun_out/ncu_mbconv python tools/run_forward.py 32 1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_strip -s 20 -c 1 -o gpurun_out/ncu_dw python tools/run_forward.py 32 1 > gpurun_out/ncu_c.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 300 python tools/bench_transformer.py cfg4 bf16 2>&1 | tail -1 > gpurun_out/bench_transformer_cfg4.json; cat gpurun_out/bench_transformer_cfg4.json | cut -c1-300
timeout 300 python tools/bench_transformer.py default bf16 2>&1 | tail -1 > gpurun_out/bench_transformer_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
