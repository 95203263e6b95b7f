set -x
CMD="python tools/profile_step.py --frames 64 --queries 500000"
$CMD > gpurun_out/plain64.log 2>&1 && \
ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f64.csv $CMD > gpurun_out/ncu64a.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:gemm_bf16_kernel -s 420 -c 4 -f -o gpurun_out/r01_gemm $CMD > gpurun_out/ncu64b.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:attn_d64_kernel -c 2 -f -o gpurun_out/r01_attn $CMD > gpurun_out/ncu64c.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:"ln_rows_kernel|ae_query_kernel|boundary_kernel" -c 3 -f -o gpurun_out/r01_misc $CMD > gpurun_out/ncu64d.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:xattn_fused_kernel -s 100 -c 2 -f -o gpurun_out/r01_xattn $CMD > gpurun_out/ncu64e.log 2>&1
tail -n 2 gpurun_out/plain64.log gpurun_out/ncu64a.log gpurun_out/ncu64b.log gpurun_out/ncu64c.log gpurun_out/ncu64d.log gpurun_out/ncu64e.log
ls -la gpurun_out | tail -12
