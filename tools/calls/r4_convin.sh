#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encoder.py -x -q -s > gpurun_out/r4_convin_test.log 2>&1; echo "rc=$?" >> gpurun_out/r4_convin_test.log
tail -3 gpurun_out/r4_convin_test.log; grep "conv_in" gpurun_out/r4_convin_test.log | head
timeout 100 python tools/gpu_time_conv_in.py 2>&1 | tee gpurun_out/r4_convin_time.log
