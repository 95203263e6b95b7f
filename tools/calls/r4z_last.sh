#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_denoiser.py -x -q -s > gpurun_out/r4z_last_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4z_last_tests.log
tail -n 3 gpurun_out/r4z_last_tests.log; grep "error /" gpurun_out/r4z_last_tests.log | sort -t= -k2 -g | tail -2
timeout 40 python tools/gpu_time_boundary.py --frames 64 2>&1 | tail -1
