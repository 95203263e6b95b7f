"""One warm-up step + one NVTX-bracketed step of the bench workload, for ncu:
    ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --frames 1
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from rald_b200 import _lib, postproc, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--queries", type=int, default=500000)
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
os.environ.setdefault("RALD_B200_GRAPH", "0")
dev = torch.device("cuda", 0)
net, vae = bench.build_models(dev)
cube = synth.radar_cube(a.frames, seed=1024).to(dev)
q = synth.query_points(1, a.queries).expand(a.frames, a.queries, 3).contiguous().to(dev)
seeds = torch.arange(a.frames)


def step():
    z = net.sample(cube, batch_seeds=seeds, cond_type="radar")
    lg = vae.decode(z, q)
    return postproc.occupied_points(lg, q, float(lg.mean()), bench.PC_RANGE, True, False, True, capacity=a.queries // 4)


for _ in range(a.warm):
    step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
torch.cuda.nvtx.range_push("profiled_step")
step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("launches in the profiled step:", _lib.launch_count() - n0)
