#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_train_ops.py > gpurun_out/r2h_plain_ops.log 2>&1 || exit 1
timeout 300 python tools/gpu_time_train_full.py 8 > gpurun_out/r2h_plain_full.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_train_b8_full.csv \
  python tools/gpu_time_train_full.py 8 > gpurun_out/r2h_ncu_full.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_train_b8_full.csv > gpurun_out/r02_launches_train_b8_full.md
gzip -f gpurun_out/r02_launches_train_b8_full.csv
ncu --set full --clock-control none --nvtx --nvtx-include "profiled/" -f -o /tmp/r02_train_ops \
  python tools/profile_train_ops.py > gpurun_out/r2h_ncu_ops.log 2>&1
python tools/ncu_summary.py /tmp/r02_train_ops.ncu-rep > gpurun_out/r02_train_ops_ncu.md
cat gpurun_out/r2h_plain_full.log | tail -n 2
head -n 30 gpurun_out/r02_train_ops_ncu.md
du -sh gpurun_out
