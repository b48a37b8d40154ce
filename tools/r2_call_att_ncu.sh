#!/bin/bash
mkdir -p gpurun_out
FTC_ATT_TC=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 20 -c 1 -o gpurun_out/r02an_att_tc -f python tools/bench_transformer.py cfg4 > gpurun_out/r02an_ncu_tc.log 2>&1
tail -n 3 gpurun_out/r02an_ncu_tc.log
