#!/bin/bash
mkdir -p gpurun_out
set -x
CMD="python tools/profile_step.py --frames 64 --queries 500000"
timeout 300 $CMD > gpurun_out/plain64.log 2>&1 && \
timeout 900 ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03_launches_f64.csv $CMD > gpurun_out/ncu64a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:gemm_bf16_kernel -s 420 -c 4 -f -o gpurun_out/r03_gemm $CMD > gpurun_out/ncu64b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:attn_d64_streams -s 20 -c 2 -f -o gpurun_out/r03_attn $CMD > gpurun_out/ncu64c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:"xattn_fused_kernel|ae_query_kernel|conv3d_kernel" -s 30 -c 4 -f -o gpurun_out/r03_misc $CMD > gpurun_out/ncu64d.log 2>&1
CMD8="python tools/profile_step.py --frames 8 --queries 500000"
timeout 300 $CMD8 > gpurun_out/plain8.log 2>&1 && \
timeout 900 ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03_launches_f8.csv $CMD8 > gpurun_out/ncu8a.log 2>&1
tail -n 2 gpurun_out/plain64.log gpurun_out/ncu64a.log gpurun_out/ncu64b.log gpurun_out/ncu64c.log gpurun_out/ncu64d.log gpurun_out/ncu8a.log
gzip -f -k gpurun_out/r03_launches_f64.csv gpurun_out/r03_launches_f8.csv
ls -la gpurun_out | tail -n 12
