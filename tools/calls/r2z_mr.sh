#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/r2z_mr.log 2>&1; tail -n 3 gpurun_out/r2z_mr.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-weak > gpurun_out/r2z_n2.json 2> gpurun_out/r2z_n2.err
python -c "
import json
l=json.loads(open('gpurun_out/r2z_n2.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, l['e2e']['value'], l.get('sharding_check',{}).get('identical'))
"
