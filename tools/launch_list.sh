#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of two B=32 bf16 forwards -> gpurun_out/launches.csv ; bench line -> gpurun_out/bench.log
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/run_forward.py 32 2 > gpurun_out/launch_list.log 2>&1
tail -2 gpurun_out/launch_list.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
