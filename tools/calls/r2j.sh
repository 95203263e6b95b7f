nvidia-smi -L | wc -l
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2j_n$N.json 2> gpurun_out/r2j_n$N.err
tail -2 gpurun_out/r2j_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_n$N.json'))
print({k: d[k] for k in ('value','ms_per_step','scaling','n_gpus','gpu_launches')}, d['e2e'], d.get('sharding_check'), d.get('weak_scaling'))
"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --frames-per-gpu 32 --steps 3 --warmup 3 --quick > gpurun_out/r2j_cfg4.json 2> gpurun_out/r2j_cfg4.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_cfg4.json'))
print({k: d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e'], d.get('sharding_check'), d['config']['workload'])
"
