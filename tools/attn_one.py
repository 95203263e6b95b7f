"""One self-attention launch at B frames (for ncu). python tools/attn_one.py [frames]"""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import _lib
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, Sq, Skv = 8, 512, 512
D = H * 64
q = torch.randn(B * Sq, D, device=dev).bfloat16(); k = torch.randn(B * Skv, D, device=dev).bfloat16()
v = torch.randn(B * Skv, D, device=dev).half(); o = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16)
args = (q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, B, H, Sq, Skv, 0.125, _lib.cur_stream())
for _ in range(3):
    _lib.call("rald_attn_d64", *args)
torch.cuda.synchronize()
print("ok")
