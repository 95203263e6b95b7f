#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; tail -n 3 gpurun_out/r2v_tests.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick"
$B > gpurun_out/r2v_f64.json 2> gpurun_out/r2v_f64.err
$B --frames-per-gpu 8 > gpurun_out/r2v_f8.json 2> gpurun_out/r2v_f8.err
$B --frames-per-gpu 1 > gpurun_out/r2v_f1.json 2> gpurun_out/r2v_f1.err
for f in gpurun_out/r2v_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['gpu_launches'])"; done
tail -n 2 gpurun_out/r2v_f64.err
timeout 300 python tools/gpu_time_train.py 8 64 --no-anchor > gpurun_out/r2v_train.log 2>&1; grep -A2 "\"batch64\"\|\"batch8\"\|batch8_with" gpurun_out/r2v_train.log
