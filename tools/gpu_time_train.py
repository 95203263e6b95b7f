"""Times the training step (bench.py's train_step leg) alone.   python tools/gpu_time_train.py [batch ...]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

batches = tuple(int(a) for a in sys.argv[1:]) or (8, 64)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
print(json.dumps(bench.leg_train_step(dev, bench.load_peaks(), batches), indent=1))
