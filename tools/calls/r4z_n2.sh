#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/r4z_mr.log 2>&1; tail -n 2 gpurun_out/r4z_mr.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --quick --no-weak > gpurun_out/r4z_n2.json 2> gpurun_out/r4z_n2.err
echo "rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r4z_n2.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],2), d.get('sharding_check'), d['scaling'])"
