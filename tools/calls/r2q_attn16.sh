#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_check_attn.py > gpurun_out/r2q_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_check.log
grep -v Warning gpurun_out/r2q_check.log | tail -n 32
