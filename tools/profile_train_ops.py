"""The training step's own kernels at their production sizes, once each after a warm-up (for `ncu --set full`):
attention backward (64 frames x 8 heads, 512 x 512), one level-0 convolution weight gradient (8 frames, 128 x 64 x 32
voxels, 64 -> 64 channels: pad + transpose passes and the multi-tap split-K GEMM), its dgrad through the forward
convolution kernel, GroupNorm + swish backward, LayerNorm backward, a denoiser weight-gradient GEMM (split-K).
    python tools/profile_train_ops.py"""
import sys

import torch
import torch.nn as nn

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from rald_b200 import _lib  # noqa: E402
from rald_b200.runtime_encoder_train import EncoderTrainRuntime, _Conv3  # noqa: E402

DEV, BF = "cuda:0", torch.bfloat16
s = _lib.cur_stream()
g = torch.Generator().manual_seed(0)


def attention_backward(frames=64, Sq=512, Skv=512):
    W = 512
    q = torch.randn(frames * Sq, W, generator=g).to(DEV).to(BF)
    k = torch.randn(frames * Skv, W, generator=g).to(DEV).to(BF)
    v16 = torch.randn(frames * Skv, W, generator=g).to(DEV).to(torch.float16)
    do = torch.randn(frames * Sq, W, generator=g).to(DEV).to(BF)
    o = torch.empty_like(q)
    stats = torch.empty(frames * Sq, 8, 2, device=DEV)
    _lib.call("rald_attn_d64_stats", q.data_ptr(), W, k.data_ptr(), W, v16.data_ptr(), W, o.data_ptr(), W, frames, 8, Sq, Skv,
              0.125, stats.data_ptr(), s)
    vb = torch.empty(frames * Skv, W, device=DEV, dtype=BF)
    _lib.call("rald_center_cast_f16_bf16", v16.data_ptr(), W, vb.data_ptr(), W, frames, Skv, W, s)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(k)
    lse = torch.empty(frames * 8 * Sq, device=DEV)
    ds = torch.empty(frames * 8 * Sq, device=DEV)
    _lib.call("rald_attn_d64_bwd", q.data_ptr(), W, k.data_ptr(), W, vb.data_ptr(), W, do.data_ptr(), W, stats.data_ptr(),
              lse.data_ptr(), ds.data_ptr(), dq.data_ptr(), W, dk.data_ptr(), W, dv.data_ptr(), W, frames, 8, Sq, Skv, 0.125, s)


def conv_backward(B=8, dims=(128, 64, 32), c=64):
    rt = EncoderTrainRuntime.__new__(EncoderTrainRuntime)
    rt.dev = torch.device(DEV)
    rt.groups, rt.eps = 32, 1e-6
    conv = nn.Conv3d(c, c, 3, padding=1).to(DEV)
    cv = _Conv3("c", conv, rt.dev)
    V = dims[0] * dims[1] * dims[2]
    x = torch.randn(B, V, c, device=DEV)
    dy = torch.randn(B, V, c, device=DEV)
    x16 = rt._cast(x)
    rt._conv3_wgrad(dy, x16, c, c, dims, 1)
    rt._conv3_dgrad(rt._dy16(dy, cv), cv, dims)
    nrm = dict(name="n", g=torch.ones(c, device=DEV), b=torch.zeros(c, device=DEV))
    _, st = rt._gn(x, nrm, 0)
    rt._gn_bwd(x, st, nrm, dy, 1, dy, {})


def dit_ops(T=32768):
    x = torch.randn(T, 512, device=DEV)
    dy = torch.randn(T, 512, device=DEV).to(BF)
    dh = torch.randn(T, 512, device=DEV)
    mod = torch.randn(T // 512, 2, 512, device=DEV)
    ws = torch.empty((T // 64) * 1024, device=DEV)
    dmod = torch.empty(T // 512, 2, 512, device=DEV)
    _lib.call("rald_ln_bwd", x.data_ptr(), dy.data_ptr(), mod.data_ptr(), 1024, 512, 1, dh.data_ptr(), 1, ws.data_ptr(),
              ws.numel(), dmod.data_ptr(), 1024, 512, 0, T, 512, 1e-5, s)
    a_t = torch.randn(512, T, device=DEV).to(BF)
    b_t = torch.randn(2048, T, device=DEV).to(BF)
    out = torch.zeros(512, 2048, device=DEV)
    _lib.call("rald_gemm_bf16_accum", a_t.data_ptr(), T, b_t.data_ptr(), T, out.data_ptr(), 2048, 512, 2048, T, s)
    part = torch.empty(T // 64, 512, device=DEV)
    o16 = torch.empty(T, 512, device=DEV, dtype=BF)
    o16t = torch.empty(512, T, device=DEV, dtype=BF)
    _lib.call("rald_cast_transpose", dh.data_ptr(), 1, 512, T, 512, o16.data_ptr(), 512, o16t.data_ptr(), T, part.data_ptr(), s)


for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("profiled")
    attention_backward()
    conv_backward()
    dit_ops()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.nvtx.range_pop()
print("done")
