"""GPU: the folded cross-attention sub-layer (csrc/xattn.cu: rald_xattn_fold + rald_xattn_fused, and its two-GEMM
small-batch form rald_xattn_split, bit-identical to the fused kernel) against an fp32 torch
restatement of CrossAttention.forward (model/models_radar_generation.py:35-76) + residual add on the same bf16-rounded
operands, and against the unfused kernel sequence it replaces. Bar: 1e-2 relative L2 on the sub-layer's update (bf16
operands, fp16 probabilities; north star's per-step latent bar), typically 3e-3."""
import os

import pytest
import torch

from helpers import build_denoiser, rel_l2
from rald_b200 import _lib, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(xn, kv, wq, wo, bias, h0, depth_idx, frames_slice, heads=8):
    """fp32: h0 + to_out(softmax(to_q(xn) k^T / 8) v) + b with k, v = this block's columns of kv."""
    T, dim = xn.shape
    F = len(frames_slice)
    M = T // F
    q = (xn.float() @ wq.float().t()).view(F, M, heads, 64).transpose(1, 2)
    kvf = kv.float().view(-1, 64, kv.shape[1])[frames_slice]
    k = kvf[:, :, depth_idx * 2 * dim: depth_idx * 2 * dim + dim].reshape(F, 64, heads, 64).transpose(1, 2)
    v = kvf[:, :, depth_idx * 2 * dim + dim: (depth_idx + 1) * 2 * dim].reshape(F, 64, heads, 64).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) * 64 ** -0.5, dim=-1) @ v
    att = att.transpose(1, 2).reshape(T, dim)
    return h0 + att @ wo.float().t() + bias


@torch.no_grad()
@pytest.mark.parametrize("frames,frame0,total", [(1, 0, 1), (2, 1, 4), (5, 0, 5), (38, 2, 41)])  # 38 frames: CTA pairs
def test_fused_cross_attention_kernel(frames, frame0, total, monkeypatch):
    if frames >= 37:
        monkeypatch.setenv("RALD_B200_XATTN_PAIR", "1")   # the optional cta_group::2 form (off by default)
    torch.manual_seed(5 + frames)
    depth, dim, M = 2, 512, 512
    bf = torch.bfloat16
    wq = (torch.randn(depth, dim, dim, device=DEV) * dim ** -0.5 * 2.0).to(bf)
    wo = (torch.randn(depth, dim, dim, device=DEV) * dim ** -0.5).to(bf)
    bias = torch.randn(depth, dim, device=DEV) * 0.1
    kv = (torch.randn(total * 64, depth * 2 * dim, device=DEV) * 1.5).to(bf)
    xn = torch.randn(frames * M, dim, device=DEV).to(bf)
    h0 = torch.randn(frames * M, dim, device=DEV)
    c = 64 ** -0.5 * 1.4426950408889634
    wq_t = (wq.float() * c).transpose(1, 2).contiguous().to(bf)
    kp = torch.empty(depth, 8, total, 64, dim, device=DEV, dtype=bf)
    vt = torch.empty(depth, 8, dim, total * 64, device=DEV, dtype=torch.float16)
    _lib.call("rald_xattn_fold", kv.data_ptr(), wq_t.data_ptr(), wo.data_ptr(), depth, total, kp.data_ptr(),
              vt.data_ptr(), _lib.cur_stream())
    # the folded operands themselves: K'[n][h][f][key][:] = c * K_h[f][key] @ Wq[h*64:(h+1)*64, :]
    kvf = kv.float().view(total, 64, depth, 2, 8, 64)
    want_kp = torch.einsum("fknhd,nhdi->nhfki", kvf[:, :, :, 0], (wq.float() * c).view(depth, 8, 64, dim))
    want_vt = torch.einsum("nohd,fknhd->nhofk", wo.float().view(depth, dim, 8, 64), kvf[:, :, :, 1])
    assert rel_l2(kp, want_kp) < 6e-3
    assert rel_l2(vt.view(depth, 8, dim, total, 64), want_vt) < 6e-3
    for n in range(depth):
        h = h0.clone()
        _lib.call("rald_xattn_fused", xn.data_ptr(), kp[n].data_ptr(), vt[n].data_ptr(), bias[n].data_ptr(),
                  h.data_ptr(), frames, M, frame0, total, _lib.cur_stream())
        ref = _reference(xn, kv, wq[n], wo[n], bias[n], h0, n, list(range(frame0, frame0 + frames)))
        err = rel_l2(h - h0, ref - h0)
        assert err < 1e-2, (n, err)
        # the two-GEMM form of the same sub-layer (small-batch path): same operands, same arithmetic order -> same bits
        h2 = h0.clone()
        probs = torch.empty(frames * M, 512, device=DEV, dtype=torch.float16)
        _lib.call("rald_xattn_split", xn.data_ptr(), kp[n].data_ptr(), vt[n].data_ptr(), bias[n].data_ptr(),
                  h2.data_ptr(), probs.data_ptr(), frames, M, frame0, total, _lib.cur_stream())
        assert rel_l2(h2 - h0, ref - h0) < 1e-2
        psum = probs.float().view(frames * M, 8, 64).sum(-1)
        assert float((psum - 1).abs().max()) < 2e-2          # every head's 64 probabilities sum to 1 (fp16)
        if frames < 37:                                       # (the cta_group::2 fused form orders its sums differently)
            assert torch.equal(h2, h), (n, float((h2 - h).abs().max()))
        # applying it twice accumulates (TMA reduce-add into the residual stream)
        _lib.call("rald_xattn_fused", xn.data_ptr(), kp[n].data_ptr(), vt[n].data_ptr(), bias[n].data_ptr(),
                  h.data_ptr(), frames, M, frame0, total, _lib.cur_stream())
        assert rel_l2(h - h0, 2 * (ref - h0)) < 1e-2


@torch.no_grad()
def test_fused_and_unfused_denoiser_agree(monkeypatch):
    """EDMPrecond.forward through the fused attn2 kernel vs the to_q GEMM -> attention -> to_out GEMM sequence it
    replaces (RALD_B200_FUSE_XATTN=0): same network, both within the parity bar of each other."""
    net = build_denoiser(device=DEV)
    cube = synth.radar_cube(3, seed=8).to(DEV)
    lat = synth.unit_latents([0, 1, 2]).to(DEV)
    sg = torch.tensor([2.5, 0.4, 30.0], device=DEV).reshape(3, 1, 1)
    monkeypatch.setenv("RALD_B200_FUSE_XATTN", "1")
    monkeypatch.setenv("RALD_B200_XATTN_SPLIT_BELOW", "0")       # the one-kernel form even for 3 frames
    a = net(lat * sg, sg, cube, cond_type="radar")
    monkeypatch.setenv("RALD_B200_XATTN_SPLIT_BELOW", "32")      # default: 3 frames take the two-GEMM form
    a2 = net(lat * sg, sg, cube, cond_type="radar")
    assert torch.equal(a, a2)             # both forms of the folded sub-layer: the same bits
    monkeypatch.setenv("RALD_B200_FUSE_XATTN", "0")
    b = net(lat * sg, sg, cube, cond_type="radar")
    assert not torch.equal(a, b)          # two different kernel sequences really ran
    assert rel_l2(a, b) < 5e-3


@torch.no_grad()
def test_fused_path_against_reference_fixtures(monkeypatch, golden):
    """The reference-fixture parity of the denoiser (tests/test_gpu_denoiser.py) with the fused attn2 kernel forced on
    for these small batches: single evaluations (shared and per-sample sigma) and the full 18-step sampler trace."""
    monkeypatch.setenv("RALD_B200_FUSE_XATTN", "1")
    monkeypatch.setenv("RALD_B200_XATTN_SPLIT_BELOW", "0")
    net = build_denoiser(device=DEV)
    g = golden("denoiser_eval")
    lat = synth.unit_latents([0, 1])
    n0 = _lib.launch_count()
    for sigma in (80.0, 1.5, 0.002):
        out = net((lat * sigma).to(DEV), torch.tensor(sigma), g["tokens2"].to(DEV), "radar")
        assert rel_l2(out, g[f"denoised_{sigma}"]) < 1e-2
    sg = torch.tensor([3.0, 0.2]).reshape(2, 1, 1)
    out = net((lat * sg).to(DEV), sg.to(DEV), g["tokens2"].to(DEV), "radar")
    assert rel_l2(out, g["denoised_per_sample"]) < 1e-2
    fused_launches = _lib.launch_count() - n0
    tokens = golden("radar_cond")["tokens_dense"].to(DEV)
    ref = golden("sampler_trace")["trace"]
    x, trace = net.sample_from_latents(synth.unit_latents([0]).to(DEV), tokens, trace=True)
    errs = [rel_l2(trace[i, 0], ref[i]) for i in range(ref.shape[0])]
    print("fused attn2, per-step rel-L2:", " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < 1e-2
    # the fused sequence really ran: 24 blocks x (3 kernels -> 1) fewer launches per evaluation, 2 x 24 x 8 + 1 fold GEMMs
    monkeypatch.setenv("RALD_B200_FUSE_XATTN", "0")
    n1 = _lib.launch_count()
    for sigma in (80.0, 1.5, 0.002):
        net((lat * sigma).to(DEV), torch.tensor(sigma), g["tokens2"].to(DEV), "radar")
    net((lat * sg).to(DEV), sg.to(DEV), g["tokens2"].to(DEV), "radar")
    unfused_launches = _lib.launch_count() - n1
    # (+ 1: the evaluation boundary's weight pack, launched once when the runtime packs the module's weights)
    assert fused_launches == unfused_launches + 4 * (2 * 24 * 8 - 2 * 24) + 1
