#!/bin/bash
# ncu launch list of the transformer cfg4 bench (3 warm-up + 10 timed forwards + predictor) -> gpurun_out/tf_launches.csv
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tf_launches.csv python tools/bench_transformer.py cfg4 bf16 > gpurun_out/tf_launch_list.log 2>&1
tail -1 gpurun_out/tf_launch_list.log
