python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; tail -15 gpurun_out/r2b_tests.log
python tools/gpu_chamfer_probe.py > gpurun_out/r2b_probe.log 2>&1; tail -20 gpurun_out/r2b_probe.log
B="python bench.py --quick --steps 3 --warmup 3"
$B --frames-per-gpu 8 > gpurun_out/r2b_f8_base.json 2> gpurun_out/r2b_f8_base.err
RALD_B200_FUSE_XATTN_MIN_FRAMES=1 $B --frames-per-gpu 8 > gpurun_out/r2b_f8_fused.json 2> gpurun_out/r2b_f8_fused.err
RALD_B200_CHAINS=2 $B --frames-per-gpu 8 > gpurun_out/r2b_f8_c2.json 2> gpurun_out/r2b_f8_c2.err
RALD_B200_CHAINS=4 $B --frames-per-gpu 8 > gpurun_out/r2b_f8_c4.json 2> gpurun_out/r2b_f8_c4.err
RALD_B200_CHAINS=2 RALD_B200_FUSE_XATTN_MIN_FRAMES=1 $B --frames-per-gpu 8 > gpurun_out/r2b_f8_c2_fused.json 2> gpurun_out/r2b_f8_c2_fused.err
$B --frames-per-gpu 16 > gpurun_out/r2b_f16_base.json 2> gpurun_out/r2b_f16_base.err
RALD_B200_FUSE_XATTN_MIN_FRAMES=1 $B --frames-per-gpu 16 > gpurun_out/r2b_f16_fused.json 2> gpurun_out/r2b_f16_fused.err
RALD_B200_CHAINS=2 $B --frames-per-gpu 16 > gpurun_out/r2b_f16_c2.json 2> gpurun_out/r2b_f16_c2.err
RALD_B200_CHAINS=2 $B --frames-per-gpu 1 > gpurun_out/r2b_f1_base.json 2> gpurun_out/r2b_f1_base.err
python bench.py > gpurun_out/r2b_full.json 2> gpurun_out/r2b_full.err
for f in gpurun_out/r2b_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1))"; done
tail -3 gpurun_out/r2b_*.err | tail -40
