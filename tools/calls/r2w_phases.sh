#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_phases.py > gpurun_out/r2w_gemm_phases.log 2>&1; grep -v Warn gpurun_out/r2w_gemm_phases.log
