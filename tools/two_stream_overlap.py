"""Experiment (B200): do two independent 32-frame micro-batches on two streams overlap usefully (LayerNorm /
boundary passes of one under the tensor-bound GEMMs of the other), compared with one 64-frame micro-batch?
Sampler only, tokens injected. python tools/two_stream_overlap.py"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from helpers import build_denoiser
from rald_b200 import synth


def timed(fn, n=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


net_a = build_denoiser(device="cuda")
net_b = build_denoiser(device="cuda")
tok = torch.randn(64, 64, 512, device="cuda")
lat = synth.unit_latents(range(64)).cuda()

os.environ["RALD_B200_MICROBATCH"] = "64"
ms64 = timed(lambda: net_a.sample_from_latents(lat, tok))
print(f"one stream, 64 frames, micro-batch 64: {ms64:.1f} ms  ({64 / ms64 * 1e3:.1f} frames/s sampler only)", flush=True)

os.environ["RALD_B200_MICROBATCH"] = "32"
ms32 = timed(lambda: net_a.sample_from_latents(lat, tok))
print(f"one stream, 64 frames, micro-batch 32: {ms32:.1f} ms", flush=True)

s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def two():
    cur = torch.cuda.current_stream()
    s0.wait_stream(cur); s1.wait_stream(cur)
    with torch.cuda.stream(s0):
        net_a.sample_from_latents(lat[:32], tok[:32])
    with torch.cuda.stream(s1):
        net_b.sample_from_latents(lat[32:], tok[32:])
    cur.wait_stream(s0); cur.wait_stream(s1)


ms2 = timed(two)
print(f"two streams, 2 x 32 frames: {ms2:.1f} ms  ({64 / ms2 * 1e3:.1f} frames/s sampler only)", flush=True)
