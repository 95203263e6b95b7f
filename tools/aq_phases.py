"""Phase stamps of the decoder-query kernel (rald_ae_query_debug_buffer): python tools/aq_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import build_ae
from rald_b200 import _lib, synth

ae = build_ae(device="cuda:0")
z = torch.randn(2, 512, 32, device="cuda:0")
q = synth.query_points(1, 500000).expand(2, 500000, 3).contiguous().cuda()
ae.decode(z, q); torch.cuda.synchronize()
dbg = torch.zeros(64, device="cuda:0", dtype=torch.int64)
_lib.call("rald_ae_query_debug_buffer", dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ae.decode(z, q); e1.record(); torch.cuda.synchronize()
_lib.call("rald_ae_query_debug_buffer", 0)
d = dbg.cpu().view(4, 16)
names = ["start", "feat", "E ready", "stats", "Qn", "S0", "S1", "S2", "S3", "done"]
t0 = int(d[0, 0])
for t in range(4):
    print(f"tile {t}: " + "  ".join(f"{n}={(int(d[t, i]) - t0) / 1e3:.2f}" for i, n in enumerate(names)))
print(f"decode of 2 x 500000 queries (stack cached): {e0.elapsed_time(e1):.3f} ms")
