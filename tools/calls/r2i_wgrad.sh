#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"; python tools/gpu_time_wgrad.py
echo "== G1"; RALD_B200_WGRAD_G1=1 python tools/gpu_time_wgrad.py
echo "== BN128"; RALD_B200_WGRAD_BN=128 python tools/gpu_time_wgrad.py
echo "== BN128 G1"; RALD_B200_WGRAD_BN=128 RALD_B200_WGRAD_G1=1 python tools/gpu_time_wgrad.py
echo "== BN64"; RALD_B200_WGRAD_BN=64 python tools/gpu_time_wgrad.py
} > gpurun_out/r2i_wgrad.log 2>&1
cat gpurun_out/r2i_wgrad.log
