#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_bwd.py -q -s --timeout 300 -k "attention or training_step" > gpurun_out/r2n_attn.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_attn.log
grep -n "attention backward\|gradients of\|passed\|failed\|   model" gpurun_out/r2n_attn.log | head -12
timeout 600 python tools/gpu_time_train.py 8 64 --no-anchor 2>&1 | grep -A2 "\"batch64\"\|\"batch8\"\|batch8_with"
