#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_time_wgrad.py > gpurun_out/r2i_wgrad.log 2>&1
cat gpurun_out/r2i_wgrad.log
