"""Dev timing (B200): rald_enc_conv_in on 32 cubes of 128 x 64 x 32 (one micro-batch of the encoder), CUDA events."""
import json, sys
sys.path.insert(0, ".")
import torch
from rald_b200 import _lib
try:
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6544.7)
except OSError:
    peak = 6544.7   # the pool's measured copy bandwidth when the driver-written file is absent
B, D, H, W = 32, 128, 64, 32
x = torch.rand(B, D, H, W, 1, device="cuda")
w = torch.randn(64, 1, 3, 3, 3, device="cuda") * 0.2
b = torch.randn(64, device="cuda")
out = torch.empty(B, D, H, W, 64, device="cuda")
def launch():
    _lib.call("rald_enc_conv_in", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, D, H, W, 1, 64, _lib.cur_stream())
for _ in range(3): launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): launch()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byts = out.numel() * 4 + x.numel() * 4
print(f"conv_in 32 x 128x64x32 -> 64 ch: {ms:.3f} ms per launch, {byts / ms / 1e6:.0f} GB/s = {byts / ms / 1e6 / peak:.3f} of {peak:.0f}")
