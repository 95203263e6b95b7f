#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -s --timeout 600 > gpurun_out/r2k_ddp.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_ddp.log
tail -15 gpurun_out/r2k_ddp.log
