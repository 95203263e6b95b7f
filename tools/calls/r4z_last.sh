#!/bin/bash
mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -q -x > gpurun_out/r4z_last_full.log 2>&1; echo "rc=$?" >> gpurun_out/r4z_last_full.log
tail -n 3 gpurun_out/r4z_last_full.log
