#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_denoiser.py tests/test_gpu_e2e_chamfer.py -q --timeout 600 > gpurun_out/r2j_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_tests.log
for t in 1 0 1 0; do
RALD_B200_GEMM_TAIL=$t timeout 600 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2j_tail$t.json 2> gpurun_out/r2j_tail$t.err
python -c "
import json; l=json.loads(open('gpurun_out/r2j_tail$t.json').read().strip().splitlines()[-1]); print('tail=$t', l['value'], l['ms_per_step'], l['e2e']['value'])"
done
