"""Per-tile phase stamps of the attention kernel (CTA 0). python tools/attn_phases.py"""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import _lib
dev = "cuda"
L = _lib.lib()
for (B, H, Sq, Skv) in [(64, 8, 512, 512), (64, 8, 512, 64), (1, 8, 512, 512)]:
    D = H * 64
    q = torch.randn(B * Sq, D, device=dev).bfloat16(); k = torch.randn(B * Skv, D, device=dev).bfloat16()
    v = torch.randn(B * Skv, D, device=dev).half(); o = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(16 * 2 * 8, dtype=torch.int64, device=dev)
    args = (q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, B, H, Sq, Skv, 0.125, _lib.cur_stream())
    _lib.call("rald_attn_d64", *args)
    torch.cuda.synchronize()
    L.rald_attn_debug_buffer(dbg.data_ptr())
    _lib.call("rald_attn_d64", *args)
    torch.cuda.synchronize()
    L.rald_attn_debug_buffer(0)
    d = dbg.view(16, 2, 8).cpu()
    t0 = d[0, 0, 0]
    print(f"B={B} Skv={Skv}: per tile/group: start S-ready P-written stats O-ready stored (us from kernel's first stamp)")
    for t in range(min(8, 16)):
        if d[t, 0, 0] == 0:
            break
        for g in range(2):
            print(f"  tile {t} g{g}: " + " ".join(f"{(int(d[t, g, i]) - int(t0)) / 1000:7.2f}" for i in range(6)))
