#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-weak > gpurun_out/r2z_n8.json 2> gpurun_out/r2z_n8.err
echo "rc=$?"; tail -n 3 gpurun_out/r2z_n8.err
python -c "
import json
l=json.loads(open('gpurun_out/r2z_n8.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches')}, l['e2e']['value'], l.get('sharding_check'))
"
