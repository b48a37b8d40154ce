#!/bin/bash
# tcgen05 attention: parity (ops test + transformer tests), then A/B of the cfg4 forward with the kernel off / on
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k attention 2>&1 | tail -15 > gpurun_out/r02ah_att_tests.txt
cat gpurun_out/r02ah_att_tests.txt
timeout 400 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_backend_abi.py -x -q -m gpu 2>&1 | tail -8 | tee -a gpurun_out/r02ah_att_tests.txt
FTC_ATT_TC=0 timeout 200 python tools/bench_transformer.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02ah_tf_cfg4_att_mma.json
FTC_ATT_TC=1 timeout 200 python tools/bench_transformer.py cfg4 2>&1 | tail -1 | tee gpurun_out/r02ah_tf_cfg4_att_tc.json
