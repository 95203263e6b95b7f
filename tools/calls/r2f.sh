./tools/micro/_bin/exp_rate > gpurun_out/r2f_exp.log 2>&1; grep "256 threads" gpurun_out/r2f_exp.log
CMD="python tools/profile_step.py --frames 8 --queries 65536"
$CMD > gpurun_out/r2f_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:attn_d64_kernel -s 10 -c 1 -f -o gpurun_out/r02_attn_f8 $CMD > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_plain.log gpurun_out/r2f_ncu.log; ls -la gpurun_out/r02_attn_f8.ncu-rep
