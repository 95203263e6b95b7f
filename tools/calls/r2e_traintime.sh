#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_time_train.py 8 64 > gpurun_out/r2e_traintime.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_traintime.log
tail -50 gpurun_out/r2e_traintime.log
