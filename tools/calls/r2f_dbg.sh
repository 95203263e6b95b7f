#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_debug_enc.py > gpurun_out/r2f_dbg.log 2>&1
tail -20 gpurun_out/r2f_dbg.log
