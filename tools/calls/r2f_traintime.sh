#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_time_train.py 8 64 > gpurun_out/r2f_traintime.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_traintime.log
tail -45 gpurun_out/r2f_traintime.log
