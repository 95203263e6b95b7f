#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2s_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2s_tests.log | head -20; tail -n 3 gpurun_out/r2s_tests.log
timeout 200 python tools/gpu_check_attn.py 2>&1 | grep -A1 "B=64 H=8 Sq=512 Skv=512\|B=8 H=8 Sq=512 Skv=512\|B=1 H=8 Sq=512 Skv=512 amp=1.0\|FAIL"
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick"
$B > gpurun_out/r2s_f64.json 2> gpurun_out/r2s_f64.err
$B --frames-per-gpu 8 > gpurun_out/r2s_f8.json 2> gpurun_out/r2s_f8.err
$B --frames-per-gpu 1 > gpurun_out/r2s_f1.json 2> gpurun_out/r2s_f1.err
RALD_B200_ATTN_STREAMS=0 $B > gpurun_out/r2s_f64_old.json 2> gpurun_out/r2s_f64_old.err
RALD_B200_ATTN_STREAMS=0 $B --frames-per-gpu 8 > gpurun_out/r2s_f8_old.json 2> gpurun_out/r2s_f8_old.err
RALD_B200_ATTN_STREAMS=0 $B --frames-per-gpu 1 > gpurun_out/r2s_f1_old.json 2> gpurun_out/r2s_f1_old.err
for f in gpurun_out/r2s_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['gpu_launches'])"; done
