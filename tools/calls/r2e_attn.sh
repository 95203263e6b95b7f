#!/bin/bash
mkdir -p gpurun_out
{
for args in "1 128 128 0" "1 128 128 1" "2 512 512 0" "3 512 64 0" "2 256 384 0"; do
  echo "=== $args"; timeout 120 python tools/gpu_check_attn_bwd.py $args 2>&1 | tail -8
done
} > gpurun_out/r2e_attn.log 2>&1
cat gpurun_out/r2e_attn.log
