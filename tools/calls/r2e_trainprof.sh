#!/bin/bash
mkdir -p gpurun_out
for B in 64 8; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2e_launches_train_b$B.csv \
  python tools/gpu_time_train.py $B --no-anchor --once > gpurun_out/r2e_trainprof_b$B.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_trainprof_b$B.log
done
tail -3 gpurun_out/r2e_trainprof_b64.log
