python -m pytest tests -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; grep -n "record\|passed\|failed\|FAILED\|Error" gpurun_out/r2e_tests.log | head -20
python tools/attn_phases.py > gpurun_out/r2e_attn_phases.log 2>&1; tail -12 gpurun_out/r2e_attn_phases.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
$B > gpurun_out/r2e_f64.json 2> gpurun_out/r2e_f64.err
$B --quick --frames-per-gpu 8 > gpurun_out/r2e_f8.json 2> gpurun_out/r2e_f8.err
$B --quick --frames-per-gpu 16 > gpurun_out/r2e_f16.json 2> gpurun_out/r2e_f16.err
for f in gpurun_out/r2e_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['gpu_launches']); print([(r['family'], round(r['achieved'],1), round(r['frac'],3), r['share_of_step']) for r in d.get('rooflines',[])])"; done
