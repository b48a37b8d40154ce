#!/bin/bash
# weight-stationary schedule on/off and BN=128 plan on/off -> per-op dumps
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/profile_ops.py 32 0.1 > gpurun_out/ab2_$name.txt 2>&1; cp gpurun_out/ops_all.txt gpurun_out/ab2_${name}_all.txt; echo "== $name: $(head -1 gpurun_out/ab2_$name.txt)"; sed -n 3,4p gpurun_out/ab2_$name.txt; }
run base A=0
run nobstat FTC_TMA_FLAGS=2048
run nobn FTC_NO_BSTAT_BN=1
run neither FTC_TMA_FLAGS=2048 FTC_NO_BSTAT_BN=1
