#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r4a_tests.log 2>&1; tail -n 2 gpurun_out/r4a_tests.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick"
$B > gpurun_out/r4a_f64.json 2> gpurun_out/r4a_f64.err
$B --frames-per-gpu 8 > gpurun_out/r4a_f8.json 2> gpurun_out/r4a_f8.err
$B --frames-per-gpu 1 > gpurun_out/r4a_f1.json 2> gpurun_out/r4a_f1.err
for f in gpurun_out/r4a_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['gpu_launches'])"; done
