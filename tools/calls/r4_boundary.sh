#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_boundary.py -x -q > gpurun_out/r4_boundary_test.log 2>&1; echo "rc=$?" >> gpurun_out/r4_boundary_test.log
tail -3 gpurun_out/r4_boundary_test.log
timeout 120 python tools/gpu_time_boundary.py 2>&1 | tee gpurun_out/r4_boundary_time.log
