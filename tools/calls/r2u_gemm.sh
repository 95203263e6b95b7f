#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_denoiser.py tests/test_gpu_attention.py tests/test_gpu_train_bwd.py -q -x > gpurun_out/r2u_tests.log 2>&1; tail -n 3 gpurun_out/r2u_tests.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick"
$B > gpurun_out/r2u_f64.json 2> gpurun_out/r2u_f64.err
$B --frames-per-gpu 8 > gpurun_out/r2u_f8.json 2> gpurun_out/r2u_f8.err
$B --frames-per-gpu 1 > gpurun_out/r2u_f1.json 2> gpurun_out/r2u_f1.err
for f in gpurun_out/r2u_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['gpu_launches'])"; done
tail -n 2 gpurun_out/r2u_f64.err
