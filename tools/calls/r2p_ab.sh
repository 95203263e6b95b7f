#!/bin/bash
mkdir -p gpurun_out
{
for pl in 32768 0 32768 0 8192 131072; do
echo "== PANEL $pl"
RALD_B200_WGRAD_PANEL=$pl timeout 300 python tools/gpu_time_train_full.py 8 6 2>&1 | tail -n 4
RALD_B200_WGRAD_PANEL=$pl timeout 100 python tools/gpu_time_wgrad.py 2>&1 | head -n 2
done
} > gpurun_out/r2p_ab.log 2>&1
cat gpurun_out/r2p_ab.log
