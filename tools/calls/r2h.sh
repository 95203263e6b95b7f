timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2h_tests.log | head -20; tail -5 gpurun_out/r2h_tests.log
timeout 120 python tools/attn_phases.py > gpurun_out/r2h_attn_phases.log 2>&1; head -10 gpurun_out/r2h_attn_phases.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
$B > gpurun_out/r2h_f64.json 2> gpurun_out/r2h_f64.err
$B --quick --frames-per-gpu 8 > gpurun_out/r2h_f8.json 2> gpurun_out/r2h_f8.err
for f in gpurun_out/r2h_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['gpu_launches']); print([(r['family'], round(r['achieved'],1), round(r['frac'],3), r['share_of_step']) for r in d.get('rooflines',[])])"; done
tail -3 gpurun_out/r2h_f64.err
