"""Issuer / stream stamps of the four-stream attention kernel (CTA 0). python tools/attn_streams_phases.py [frames]"""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import _lib
dev = "cuda"
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, Sq, Skv = 8, 512, 512
D = H * 64
q = torch.randn(B * Sq, D, device=dev).bfloat16(); k = torch.randn(B * Skv, D, device=dev).bfloat16()
v = torch.randn(B * Skv, D, device=dev).half(); o = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16)
dbg = torch.zeros(256 + 4 * 32 * 3, dtype=torch.int64, device=dev)
args = (q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, B, H, Sq, Skv, 0.125, _lib.cur_stream())
_lib.call("rald_attn_d64", *args)
torch.cuda.synchronize()
L.rald_attn_streams_debug_buffer(dbg.data_ptr())
_lib.call("rald_attn_d64", *args)
torch.cuda.synchronize()
L.rald_attn_streams_debug_buffer(0)
d = dbg.cpu()
iss = d[:256].view(32, 4, 2); st = d[256:].view(4, 32, 3)
t0 = int(iss[0, 0, 1])
f = lambda x: f"{(int(x) - t0) / 1000:6.2f}" if int(x) else "   -  "
print("chunk n | per stream: issuer [PV issued, S issued] | stream [S seen, max done, P written]  (us)")
for n in range(12):
    print(f"n={n:2d} " + " | ".join(f"s{s}: {f(iss[n, s, 0])} {f(iss[n, s, 1])} / {f(st[s, n, 0])} {f(st[s, n, 1])} {f(st[s, n, 2])}" for s in range(4)))
