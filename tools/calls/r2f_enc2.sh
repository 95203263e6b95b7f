#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_encoder.py -q -s --timeout 300 > gpurun_out/r2f_enc.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_enc.log
grep -n "encoder gradients\|encoder output\|passed\|failed\|Error\|assert" gpurun_out/r2f_enc.log | head -30
timeout 900 python tools/gpu_time_train.py 8 --no-anchor > gpurun_out/r2f_traintime2.log 2>&1
grep -A3 "batch8_with_encoder\|\"batch8\"" gpurun_out/r2f_traintime2.log
