#!/bin/bash
# training-step parity tests (new backward kernels) on one B200
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_bwd.py -q -s --timeout 300 > gpurun_out/r2e_train.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_train.log
tail -40 gpurun_out/r2e_train.log
