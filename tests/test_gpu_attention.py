"""GPU: rald_attn_d64 / rald_attn_d64_stats (csrc/attn.cu for Skv < 512, csrc/attn_streams.cu — four concurrent softmax
streams with a lazily raised shift — for Skv = 512) against an fp32 torch restatement of the einsum / softmax / einsum of
CrossAttention.forward (model/models_radar_generation.py:66-75) on the same bf16 Q / K and fp16 V. Bar: 1e-2 relative L2
(bf16 operands, fp16 probabilities), typically 1.7e-3. Also: a frame computes bit-identical values alone and inside a
batch (the scheduler picks other tile groupings for other batch sizes), the rescale path of the lazy shift (key norms
growing with the key index), and the statistics contract of the training forward (exact row maximum, matching sum)."""
import pytest
import torch

from helpers import rel_l2
from rald_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALE = 0.125


def _inputs(B, H, Sq, Skv, amp=1.0, ramp=0.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    D = H * 64
    q = (torch.randn(B * Sq, D, device=DEV, generator=g) * amp).bfloat16()
    k = torch.randn(B * Skv, D, device=DEV, generator=g) * amp
    if ramp:   # later key chunks raise the running maximum by far more than 2^8
        k = (k.view(B, Skv, D) * torch.linspace(0.05, ramp, Skv, device=DEV)[None, :, None]).reshape(B * Skv, D)
    k = k.bfloat16()
    v = torch.randn(B * Skv, D, device=DEV, generator=g).half()
    return q, k, v


def _run(q, k, v, B, H, Sq, Skv, stats=False):
    D = H * 64
    o = torch.zeros(B * Sq, D, device=DEV, dtype=torch.bfloat16)
    args = [q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, B, H, Sq, Skv, SCALE]
    if stats:
        st = torch.zeros(B * Sq, H, 2, device=DEV)
        _lib.call("rald_attn_d64_stats", *args, st.data_ptr(), _lib.cur_stream())
        torch.cuda.synchronize()
        return o, st
    _lib.call("rald_attn_d64", *args, _lib.cur_stream())
    torch.cuda.synchronize()
    return o


def _scores(q, k, B, H, Sq, Skv):
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2)
    kf = k.float().view(B, Skv, H, 64).transpose(1, 2)
    return qf @ kf.transpose(-1, -2) * SCALE


def _reference(q, k, v, B, H, Sq, Skv):
    vf = v.float().view(B, Skv, H, 64).transpose(1, 2)
    out = torch.softmax(_scores(q, k, B, H, Sq, Skv), -1) @ vf
    return out.transpose(1, 2).reshape(B * Sq, H * 64)


@torch.no_grad()
@pytest.mark.parametrize("B,H,Sq,Skv,amp,ramp", [
    (1, 1, 128, 64, 1.0, 0.0), (2, 3, 256, 384, 1.0, 0.0), (5, 8, 512, 256, 1.0, 0.0),
    (1, 1, 128, 512, 1.0, 0.0),      # one tile, two streams
    (1, 8, 512, 512, 1.0, 0.0),      # batch 1 of the denoiser
    (3, 2, 384, 512, 1.0, 0.0),      # odd tile count per head: one tile per item
    (3, 8, 512, 512, 3.0, 0.0), (2, 8, 512, 512, 5.0, 0.0),
    (2, 8, 512, 512, 2.0, 4.0), (3, 2, 256, 512, 1.0, 8.0),   # lazy-shift rescale
    (40, 8, 512, 512, 1.0, 0.0),     # more items than SMs: persistent loop, K / V chunk recycling, Q slot reuse
])
def test_attention_against_fp32(B, H, Sq, Skv, amp, ramp):
    q, k, v = _inputs(B, H, Sq, Skv, amp, ramp, seed=B * 7 + Skv)
    o = _run(q, k, v, B, H, Sq, Skv)
    assert rel_l2(o, _reference(q, k, v, B, H, Sq, Skv)) < 1e-2


@torch.no_grad()
def test_frame_alone_equals_frame_in_batch():
    B, H, Sq, Skv = 24, 8, 512, 512
    q, k, v = _inputs(B, H, Sq, Skv, 2.0, 3.0, seed=11)
    o = _run(q, k, v, B, H, Sq, Skv)
    for f in (0, 7, 23):
        o1 = _run(q[f * Sq:(f + 1) * Sq].contiguous(), k[f * Skv:(f + 1) * Skv].contiguous(),
                  v[f * Skv:(f + 1) * Skv].contiguous(), 1, H, Sq, Skv)
        assert torch.equal(o1, o[f * Sq:(f + 1) * Sq])


@torch.no_grad()
@pytest.mark.parametrize("Skv,ramp", [(512, 0.0), (512, 6.0), (128, 0.0)])
def test_statistics_contract(Skv, ramp):
    """stats = (exact row maximum in log2 units with the scale folded in, sum of 2^(s - max)): rald_attn_d64_bwd and the
    chunk merge of rald_attn_d64_long recompute probabilities from them."""
    B, H, Sq = 3, 4, 256
    q, k, v = _inputs(B, H, Sq, Skv, 1.5, ramp, seed=3)
    o, st = _run(q, k, v, B, H, Sq, Skv, stats=True)
    s2 = _scores(q, k, B, H, Sq, Skv) * 1.4426950408889634          # [B, H, Sq, Skv] in log2 units
    m = s2.max(-1).values
    l = torch.exp2(s2 - m[..., None]).sum(-1)
    m_k = st[..., 0].view(B, Sq, H).permute(0, 2, 1)
    l_k = st[..., 1].view(B, Sq, H).permute(0, 2, 1)
    assert float((m_k - m).abs().max()) <= 2e-3 * float(m.abs().max()) + 1e-4
    assert rel_l2(l_k, l) < 5e-3
    assert rel_l2(o, _reference(q, k, v, B, H, Sq, Skv)) < 1e-2
