#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_check_attn.py > gpurun_out/r2q_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_check.log
grep -v Warning gpurun_out/r2q_check.log | grep -v "^ln\|Skv=64 \|Skv=128\|Skv=256\|Skv=384" | tail -n 20
timeout 300 python -m pytest tests/test_gpu_attention.py -q 2>&1 | tail -n 3
for sk in 0 1; do
echo "== skew $sk"
RALD_B200_ATTN_SKEW=$sk timeout 200 python tools/gpu_check_attn.py 2>&1 | grep -v Warning | grep -A1 "B=64 H=8 Sq=512 Skv=512\|B=8 H=8 Sq=512 Skv=512\|FAIL"
done
timeout 120 python tools/attn_streams_phases.py 64 > gpurun_out/r2r_phases.log 2>&1; cat gpurun_out/r2r_phases.log | grep -v Warn | head -n 14
