"""CPU probe behind the conditioning note of csrc/attn_bwd.cu: which rounding put 4 - 15 % of noise on the self-attention
to_q / to_k weight gradients of the deep blocks (random init) in the first version of the attention backward kernel?
The oracle (fp32, torch autograd) is run on the training fixture's inputs with ONE change at a time:
  1. q and k rounded to bf16 before the logits (straight-through), forward and backward  -> harmless (1.4e-3);
  2. attention backward by the flash-attention formulas with dO, V and O rounded to bf16 and D = rowsum(dO o O)
     ("kernel_v1")                                                                        -> reproduces it (7 - 9 %);
  3. the same with V centred per (frame, head) before rounding and D = sum_j p_ij dP_ij from the same dP ("centered",
     what the kernel does now)                                                            -> 9e-4.
Reason: dP_ij and D_i share the term dO_i . vbar (vbar = what all values of the head have in common; the deep blocks'
tokens are nearly collinear, |vbar| >> spread) and only their difference enters dS.

    python tools/probe_train_precision.py          # ~4 min on 8 cores, no GPU, no reference tree
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_denoiser, cpu_state_dict, rel_l2  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402


def grads(sd, cube, y, sigma, noise):
    for v in sd.values():
        v.grad = None
    tok = orc.process_radar_cond(sd, cube)
    D = orc.edm_precond(sd, y + noise * sigma, sigma, tok)
    weight = (sigma ** 2 + 1.0) / sigma ** 2
    loss = (weight * (D - y) ** 2).mean()
    loss.backward()
    return float(loss), {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}


def bf(x):
    return x.to(torch.bfloat16).float()


class FlashStyleAttention(torch.autograd.Function):
    """fp32 forward; backward by the formulas of csrc/attn_bwd.cu with selectable roundings:
       mode "kernel_v1": dO, V, O rounded to bf16, D = rowsum(dO o O)          (first version of the kernel)
       mode "centered":  V centred per (frame, head) before the bf16 rounding, D = sum_j p_ij dP_ij from the same dP"""
    mode = "kernel_v1"

    @staticmethod
    def forward(ctx, qh, kh, vh):
        d = qh.shape[-1]
        p = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1)
        o = p @ vh
        ctx.save_for_backward(qh, kh, vh, p, o)
        return o

    @staticmethod
    def backward(ctx, do):
        qh, kh, vh, p, o = ctx.saved_tensors
        d = qh.shape[-1]
        do_r = bf(do)
        if FlashStyleAttention.mode == "kernel_v1":
            dp = do_r @ bf(vh).transpose(-1, -2)
            D = (do_r * bf(o)).sum(-1, keepdim=True)
        else:
            vc = bf(vh - vh.mean(dim=-2, keepdim=True))
            dp = do_r @ vc.transpose(-1, -2)
            D = (p * dp).sum(-1, keepdim=True)
        ds = p * (dp - D) * d ** -0.5
        return ds @ kh, ds.transpose(-1, -2) @ qh, p.transpose(-1, -2) @ do_r


def flash_heads(q, k, v, heads):
    B, Sq, Dm = q.shape
    d = Dm // heads
    qh = q.view(B, Sq, heads, d).transpose(1, 2)
    kh = k.view(B, -1, heads, d).transpose(1, 2)
    vh = v.view(B, -1, heads, d).transpose(1, 2)
    return FlashStyleAttention.apply(qh, kh, vh).transpose(1, 2).reshape(B, Sq, Dm)


def report(tag, g1, g0):
    rows = sorted(((rel_l2(g1[k], g0[k]), k) for k in g0), reverse=True)
    print(f"{tag}: worst 6 of {len(rows)} (rel-L2 vs fp32 autograd)")
    for e, k in rows[:6]:
        print(f"   {k}: {e:.3e}")
    qk = [e for e, k in rows if k.endswith("attn1.to_q.weight") or k.endswith("attn1.to_k.weight")]
    other = [e for e, k in rows if not (k.endswith("attn1.to_q.weight") or k.endswith("attn1.to_k.weight"))]
    print(f"   attn1.to_q / to_k: max {max(qk):.3e}, median {sorted(qk)[len(qk) // 2]:.3e}; all other tensors: max "
          f"{max(other):.3e}, median {sorted(other)[len(other) // 2]:.3e}", flush=True)


def main():
    torch.set_num_threads(os.cpu_count())
    fx = np.load(os.path.join(ROOT, "tests", "golden", "train_grads.npz"))
    sd = cpu_state_dict(build_denoiser())
    for k, v in sd.items():
        if not k.startswith("radar_enc."):
            v.requires_grad_(True)
    cube = synth.radar_cube(2, seed=1024)
    y, sigma, noise = (torch.from_numpy(fx[k]) for k in ("y", "sigma", "noise"))
    loss0, g0 = grads(sd, cube, y, sigma, noise)

    plain = orc._heads_attention

    def rounded_qk(q, k, v, heads):
        rq = q + (q.to(torch.bfloat16).float() - q).detach()
        rk = k + (k.to(torch.bfloat16).float() - k).detach()
        return plain(rq, rk, v, heads)
    orc._heads_attention = rounded_qk
    try:
        loss1, g1 = grads(sd, cube, y, sigma, noise)
    finally:
        orc._heads_attention = plain
    print(f"loss fp32 {loss0:.6f}, with bf16-rounded q / k {loss1:.6f} ({abs(loss1 - loss0) / loss0:.1e})")
    report("q and k rounded to bf16 (forward and backward)", g1, g0)
    for mode in ("kernel_v1", "centered"):
        FlashStyleAttention.mode = mode
        orc._heads_attention = flash_heads
        try:
            _, g2 = grads(sd, cube, y, sigma, noise)
        finally:
            orc._heads_attention = plain
        report(f"flash-style backward, mode {mode}", g2, g0)


if __name__ == "__main__":
    main()
