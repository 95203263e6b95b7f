#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_bwd.py -q -s --timeout 300 -k "training_step or attention" > gpurun_out/r2e_train2.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_train2.log
grep -n "attention backward\|training step\|gradients of\|   model\|   radar\|passed\|failed" gpurun_out/r2e_train2.log
