"""One warm-up + one NVTX-bracketed VecSet encode (FPS / `point` variant and the default `mix` variant) of a 10 000-point
frame, for ncu:
    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "profiled_encode/" \
        -k regex:"fps_kernel|point_features_kernel|softmax_rows_kernel|posterior_kernel" -o gpurun_out/r02_encode \
        python tools/profile_encode.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import models_ae, synth  # noqa: E402

dev = torch.device("cuda", 0)
mods = []
for name in ("kl_d512_m512_l32", "kl_d512_m512_l32_mix"):
    torch.manual_seed(1024)
    mods.append(models_ae.__dict__[name](N=10000).eval().to(dev))
pc = synth.lidar_points(1, 10000, seed=1024).to(dev)
noise = synth.posterior_noise(1).to(dev)
for m in mods:
    m._runtime().encode(pc, noise)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled_encode")
for m in mods:
    m._runtime().encode(pc, noise)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("encode profiled")
