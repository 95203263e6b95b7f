#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_encoder.py -q -s --timeout 300 > gpurun_out/r2p_enc.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_enc.log
grep -n "conv3d backward\|encoder gradients\|passed\|failed\|Error\|assert" gpurun_out/r2p_enc.log | head -30
timeout 300 python tools/gpu_time_train_full.py 8 > gpurun_out/r2p_full.log 2>&1
tail -n 8 gpurun_out/r2p_full.log
