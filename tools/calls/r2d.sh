./tools/micro/_bin/exp_rate > gpurun_out/r2d_exp.log 2>&1; cat gpurun_out/r2d_exp.log
python -m pytest tests -m gpu -q > gpurun_out/r2d_tests.log 2>&1; grep -n "record\|passed\|failed\|FAILED\|Error" gpurun_out/r2d_tests.log | head -40
python - > gpurun_out/r2d_fps.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
from rald_b200 import models_ae, synth, _lib
dev = torch.device("cuda", 0)
torch.manual_seed(1024)
vae = models_ae.kl_d512_m512_l32(N=10000).eval().to(dev)
rt = vae._runtime()
for B in (1, 64, 148):
    pc = synth.lidar_points(B, 10000, seed=1).to(dev)
    ms = bench.cuda_time(lambda: rt.fps(pc, 512), 10)
    print(f"fps B={B}: {ms*1e3:.1f} us per launch, {ms*1e3/511:.3f} us per pick")
PY
cat gpurun_out/r2d_fps.log
B="python bench.py --quick --steps 3 --warmup 3"
for F in 1 8 16 32 64; do $B --frames-per-gpu $F > gpurun_out/r2d_f$F.json 2> gpurun_out/r2d_f$F.err; done
RALD_B200_XATTN_SPLIT_BELOW=0 $B --frames-per-gpu 16 > gpurun_out/r2d_f16_fusedonly.json 2> gpurun_out/r2d_f16_fusedonly.err
RALD_B200_XATTN_SPLIT_BELOW=64 $B --frames-per-gpu 32 > gpurun_out/r2d_f32_split.json 2> gpurun_out/r2d_f32_split.err
RALD_B200_FUSE_XATTN=0 $B --frames-per-gpu 1 > gpurun_out/r2d_f1_unfused.json 2> gpurun_out/r2d_f1_unfused.err
for f in gpurun_out/r2d_f*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['e2e']['value'],1), d['gpu_launches'])"; done
tail -n 3 gpurun_out/r2d_f*.err | tail -30
