"""One warm-up + one timed full training step (encoder trainable) at batch B — for ncu launch lists.
    python tools/gpu_time_train_full.py [B [steps]]"""
import sys
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from rald_b200 import synth  # noqa: E402
from rald_b200.models_radar_generation import EDMLoss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N_STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
net, _ = bench.build_models(dev)
net.train()
crit = EDMLoss()
cubes = bench.frame_cubes(0, B).to(dev)
y = (synth.unit_latents(range(B)) * 0.7).to(dev)
for it in range(N_STEPS):
    for p in net.parameters():
        p.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    crit(net, y, cubes, "radar").backward()
    e1.record()
    torch.cuda.synchronize()
    print("step", it, e0.elapsed_time(e1), "ms", flush=True)
