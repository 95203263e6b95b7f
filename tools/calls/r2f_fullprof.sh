#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2f_launches_trainfull_b8.csv \
  python tools/gpu_time_train_full.py 8 > gpurun_out/r2f_fullprof.log 2>&1
tail -3 gpurun_out/r2f_fullprof.log
