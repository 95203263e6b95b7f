# Shorter form of profile_round.sh: launch list of the bench step (+ optionally --set full of the four GEMMs of one
# denoiser block with FULL=1). Run under gpurun; the plain run must exit 0 first.
set -x
CMD="python tools/profile_step.py --frames 64 --queries 500000"
$CMD > gpurun_out/plain64.log 2>&1 && \
ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f64.csv $CMD > gpurun_out/ncu64a.log 2>&1
if [ "${FULL:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:gemm_bf16_kernel -s 420 -c 4 -f -o gpurun_out/r01_gemm $CMD > gpurun_out/ncu64b.log 2>&1
fi
tail -n 2 gpurun_out/plain64.log gpurun_out/ncu64a.log
