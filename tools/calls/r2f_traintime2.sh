#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_time_train.py 8 --no-anchor > gpurun_out/r2f_traintime2.log 2>&1
grep -A3 "batch8_with_encoder\|\"batch8\"" gpurun_out/r2f_traintime2.log
