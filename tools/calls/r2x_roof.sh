#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2x_f64.json 2> gpurun_out/r2x_f64.err
python -c "
import json
d=json.load(open('gpurun_out/r2x_f64.json'))
print(round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['gpu_launches'])
for r in d.get('rooflines',[]): print(r['family'], round(r['achieved'],1), round(r['frac'],3), r['share_of_step'])
print(d.get('latency_b1',{}).get('ms_per_frame'))
print(json.dumps(d.get('kernel_breakdown'),indent=0)[:1500])
"
tail -n 2 gpurun_out/r2x_f64.err
