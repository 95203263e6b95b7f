#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_pitch_probe.py > gpurun_out/r2o_pitch.log 2>&1; cat gpurun_out/r2o_pitch.log
