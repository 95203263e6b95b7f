python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2g_tests.log | head -20
python tools/attn_phases.py > gpurun_out/r2g_attn_phases.log 2>&1; head -12 gpurun_out/r2g_attn_phases.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
$B > gpurun_out/r2g_f64.json 2> gpurun_out/r2g_f64.err
$B --quick --frames-per-gpu 8 > gpurun_out/r2g_f8.json 2> gpurun_out/r2g_f8.err
for f in gpurun_out/r2g_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['gpu_launches']); print([(r['family'], round(r['achieved'],1), round(r['frac'],3), r['share_of_step']) for r in d.get('rooflines',[])])"; done
CMD="python tools/profile_step.py --frames 64 --queries 500000"
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:xattn_fused_kernel -s 100 -c 1 -f -o gpurun_out/r02_xattn_f64 $CMD > gpurun_out/r2g_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:ae_query_kernel -c 1 -f -o gpurun_out/r02_aequery_f64 $CMD > gpurun_out/r2g_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:attn_d64_kernel -s 100 -c 1 -f -o gpurun_out/r02_attn_f64 $CMD > gpurun_out/r2g_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_step/" -k regex:gemm_bf16_kernel -s 420 -c 4 -f -o gpurun_out/r02_gemm_f64 $CMD > gpurun_out/r2g_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
