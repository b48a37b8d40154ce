#!/bin/bash
# ncu launch list of ONE transformer cfg4 forward (cudaProfilerStart / Stop around it) -> gpurun_out/tf_launches.csv
mkdir -p gpurun_out
FTC_PROFILE_ONE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/tf_launches.csv python tools/bench_transformer.py cfg4 bf16 > gpurun_out/tf_launch_list.log 2>&1
tail -n 1 gpurun_out/tf_launch_list.log
