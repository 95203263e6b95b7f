"""Times the training step (bench.py's train_step leg) alone.   python tools/gpu_time_train.py [batch ...]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

anchors = "--no-anchor" not in sys.argv
one = "--once" in sys.argv          # one timed step after one warm-up (for ncu launch lists)
batches = tuple(int(a) for a in sys.argv[1:] if not a.startswith("--")) or (8, 64)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
print(json.dumps(bench.leg_train_step(dev, bench.load_peaks(), batches, anchors=anchors, reps=1 if one else 3), indent=1))
