#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_bwd.py -q -s --timeout 300 > gpurun_out/r2e_train.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_train.log
grep -n "training step\|gradients of\|passed\|failed\|EDMLoss over\|Error\|error" gpurun_out/r2e_train.log | head -20
timeout 600 python tools/gpu_time_train.py 8 64 --no-anchor > gpurun_out/r2e_traintime.log 2>&1
tail -22 gpurun_out/r2e_traintime.log
