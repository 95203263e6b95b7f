#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/attn_one.py 64 > gpurun_out/r2t_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_d64_streams -s 1 -c 1 -f -o gpurun_out/r02_attn_streams python tools/attn_one.py 64 > gpurun_out/r2t_ncu.log 2>&1
tail -n 3 gpurun_out/r2t_plain.log gpurun_out/r2t_ncu.log; ls -la gpurun_out/r02_attn_streams.ncu-rep
