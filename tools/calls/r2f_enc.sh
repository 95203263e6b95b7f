#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_encoder.py tests/test_gpu_train_bwd.py -q -s --timeout 300 > gpurun_out/r2f_enc.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_enc.log
grep -n "conv3d backward\|encoder gradients\|encoder output\|passed\|failed\|Error\|assert\|training step\|gradients of" gpurun_out/r2f_enc.log | head -40
