"""Runs one attention forward + backward case against torch autograd (debugging aid for csrc/attn_bwd.cu).
    python tools/gpu_check_attn_bwd.py <frames> <Sq> <Skv> <common: magnitude of a component shared by all values>"""
import sys

import torch

sys.path.insert(0, ".")
from rald_b200 import _lib  # noqa: E402

frames, Sq, Skv = (int(a) for a in sys.argv[1:4])
common = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
DEV, BF = "cuda:0", torch.bfloat16
heads, d = 8, 64
W = heads * d
g = torch.Generator().manual_seed(1)
q = torch.randn(frames * Sq, W, generator=g).to(DEV).to(BF)
k = torch.randn(frames * Skv, W, generator=g).to(DEV).to(BF)
v16 = (torch.randn(frames * Skv, W, generator=g) + common * torch.randn(1, W, generator=g)).to(DEV).to(torch.float16)
do = torch.randn(frames * Sq, W, generator=g).to(DEV).to(BF)
scale = d ** -0.5
o = torch.empty(frames * Sq, W, device=DEV, dtype=BF)
stats = torch.empty(frames * Sq, heads, 2, device=DEV)
s = _lib.cur_stream()
_lib.call("rald_attn_d64_stats", q.data_ptr(), W, k.data_ptr(), W, v16.data_ptr(), W, o.data_ptr(), W, frames, heads, Sq,
          Skv, scale, stats.data_ptr(), s)
torch.cuda.synchronize()
print("forward ok", flush=True)
vb = torch.empty(frames * Skv, W, device=DEV, dtype=BF)
_lib.call("rald_center_cast_f16_bf16", v16.data_ptr(), W, vb.data_ptr(), W, frames, Skv, W, s)
dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(k)
lse = torch.empty(frames * heads * Sq, device=DEV)
ds = torch.empty(frames * heads * Sq, device=DEV)
_lib.call("rald_attn_d64_bwd", q.data_ptr(), W, k.data_ptr(), W, vb.data_ptr(), W, do.data_ptr(), W, stats.data_ptr(),
          lse.data_ptr(), ds.data_ptr(), dq.data_ptr(), W, dk.data_ptr(), W, dv.data_ptr(), W, frames, heads, Sq, Skv, scale,
          s)
torch.cuda.synchronize()
print("backward ran", flush=True)
qf = q.float().view(frames, Sq, heads, d).transpose(1, 2).requires_grad_(True)
kf = k.float().view(frames, Skv, heads, d).transpose(1, 2).requires_grad_(True)
vf = v16.float().view(frames, Skv, heads, d).transpose(1, 2).requires_grad_(True)
p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
of = p @ vf
of.backward(do.float().view(frames, Sq, heads, d).transpose(1, 2))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


back = lambda t, S: t.transpose(1, 2).reshape(frames * S, W)
lse_ref = torch.logsumexp(qf @ kf.transpose(-1, -2) * scale, dim=-1) * 1.4426950408889634     # [f, h, Sq]
print("lse2", rel(lse.view(frames, heads, Sq), lse_ref.detach()))
print("o", rel(o, back(of, Sq)), "dq", rel(dq, back(qf.grad, Sq)), "dk", rel(dk, back(kf.grad, Skv)), "dv",
      rel(dv, back(vf.grad, Skv)))
