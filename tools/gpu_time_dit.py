"""Dev timing of the fused sampler on a B200 (tokens injected; encoder not included)."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from helpers import build_denoiser
from rald_b200 import synth

net = build_denoiser(device="cuda")
for B, mb in [(1, 1), (8, 8), (16, 16), (64, 8), (64, 16), (64, 32), (64, 64)]:
    os.environ["RALD_B200_MICROBATCH"] = str(mb)
    tok = torch.randn(B, 64, 512, device="cuda")
    lat = synth.unit_latents(range(B)).cuda()
    for _ in range(2):
        net.sample_from_latents(lat, tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    n = 3
    for _ in range(n):
        net.sample_from_latents(lat, tok)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = B * (35 * 130.494e9 + 1.686e9)
    print(f"B={B} microbatch={mb}: {ms:.1f} ms/sample-call  {B/ms*1e3:.1f} frames/s  {fl/ms/1e9:.0f} TFLOP/s  "
          f"(wall {1e3*(time.time()-t0)/n:.1f} ms)", flush=True)
