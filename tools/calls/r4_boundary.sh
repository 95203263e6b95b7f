#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_boundary.py -x -q -s > gpurun_out/r4_boundary_test.log 2>&1; echo "rc=$?" >> gpurun_out/r4_boundary_test.log
tail -4 gpurun_out/r4_boundary_test.log; grep "error /" gpurun_out/r4_boundary_test.log | sort -t= -k2 -g | tail -2
timeout 120 python tools/gpu_time_boundary.py > gpurun_out/r4_boundary_time.log 2>&1
cat gpurun_out/r4_boundary_time.log
timeout 200 python -m pytest tests/test_gpu_denoiser.py tests/test_gpu_variants.py -x -q 2>&1 | tail -5
timeout 200 ncu --set full --clock-control none --import-source on -k regex:boundary_kernel -s 10 -c 2 -f -o gpurun_out/r4_boundary python tools/gpu_time_boundary.py --frames 64 --plain > gpurun_out/r4_boundary_ncu.log 2>&1
tail -2 gpurun_out/r4_boundary_ncu.log
