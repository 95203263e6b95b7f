./tools/micro/_bin/tmem_rate > gpurun_out/r2c_tmem.log 2>&1; cat gpurun_out/r2c_tmem.log
python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; grep -n "record\|shared threshold\|common-mode\|precise sampler\|passed\|failed\|FAILED" gpurun_out/r2c_tests.log | head -40
python tools/profile_encode.py > gpurun_out/r2c_enc_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_encode/" -k regex:"fps_kernel|point_features_kernel|softmax_rows_kernel|posterior_kernel" -f -o gpurun_out/r02_encode python tools/profile_encode.py > gpurun_out/r2c_enc_ncu.log 2>&1
ncu --nvtx --nvtx-include "profiled_encode/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_encode.csv python tools/profile_encode.py > gpurun_out/r2c_enc_ncu2.log 2>&1
tail -2 gpurun_out/r2c_enc_plain.log gpurun_out/r2c_enc_ncu.log gpurun_out/r2c_enc_ncu2.log
B="python bench.py --quick --steps 3 --warmup 3"
RALD_B200_FUSE_XATTN_MIN_FRAMES=1 $B --frames-per-gpu 1 > gpurun_out/r2c_f1_fused.json 2> gpurun_out/r2c_f1_fused.err
RALD_B200_DIT_PRECISE=1 $B > gpurun_out/r2c_f64_precise.json 2> gpurun_out/r2c_f64_precise.err
RALD_B200_AE_PRECISE=0 $B > gpurun_out/r2c_f64_aeplain.json 2> gpurun_out/r2c_f64_aeplain.err
$B > gpurun_out/r2c_f64.json 2> gpurun_out/r2c_f64.err
for f in gpurun_out/r2c_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['e2e']['d2h_bytes_per_step'])"; done
