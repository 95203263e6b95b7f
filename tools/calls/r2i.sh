nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -x > gpurun_out/r2i_tests.log 2>&1; tail -5 gpurun_out/r2i_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_n2.json 2> gpurun_out/r2i_n2.err
tail -3 gpurun_out/r2i_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','scaling','n_gpus','gpu_launches')}, d['e2e'], d.get('sharding_check'), d.get('weak_scaling'))
print(d['config'])
"
