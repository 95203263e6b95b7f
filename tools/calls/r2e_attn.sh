#!/bin/bash
mkdir -p gpurun_out
{
for args in "1 128 128" "2 512 512" "3 512 64" "2 512 512 30"; do
  echo "=== $args"; timeout 120 python tools/gpu_check_attn_bwd.py $args 2>&1 | grep -v Warning | tail -4
done
} > gpurun_out/r2e_attn.log 2>&1
cat gpurun_out/r2e_attn.log
timeout 900 python -m pytest tests/test_gpu_train_bwd.py -q -s --timeout 300 > gpurun_out/r2e_train.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_train.log
grep -n "attention backward\|training step\|gradients of\|   model\|   radar\|   worst\|passed\|failed\|EDMLoss over" gpurun_out/r2e_train.log
