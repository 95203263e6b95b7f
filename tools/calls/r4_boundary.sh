#!/bin/bash
mkdir -p gpurun_out
for G in 4 3 4 3; do
echo "groups=$G"
RALD_B200_BOUNDARY_GROUPS=$G timeout 120 python tools/gpu_time_boundary.py 2>&1 | tee -a gpurun_out/r4_boundary_time_g$G.log
done
