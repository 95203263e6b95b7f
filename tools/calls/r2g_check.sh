#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_bwd.py tests/test_gpu_train_encoder.py tests/test_gpu_gemm.py tests/test_gpu_denoiser.py -q --timeout 300 > gpurun_out/r2g_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_check.log
tail -3 gpurun_out/r2g_check.log
timeout 900 python bench.py --frames-per-gpu 1 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick > gpurun_out/r2g_f1.json 2> gpurun_out/r2g_f1.err
python -c "
import json; l=json.loads(open('gpurun_out/r2g_f1.json').read().strip().splitlines()[-1]); print('batch1 ms/step', l['ms_per_step'], l['value'])"
