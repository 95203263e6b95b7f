"""GPU: long-context cross-attention (rald_attn_d64_long: key chunks of 512 + exact merge) against fp32 torch, and
the encoder-less denoiser variant (`use_radar_enc: false`, 2048 raw radar-cube tokens as context) against the fixture
of the unmodified reference. Bars: attention output 1e-2 relative L2 (bf16 operands, fp16 probabilities); denoised
output 1e-2 (north star's latent bar)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import rel_l2
from rald_b200 import _lib, synth
from test_cpu_noenc_oracle import build_noenc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@torch.no_grad()
@pytest.mark.parametrize("frames,skv", [(1, 2048), (3, 1024), (2, 1536)])
def test_long_context_attention(frames, skv):
    torch.manual_seed(frames * 7 + skv)
    heads, sq, d = 8, 256, 64
    q = torch.randn(frames * sq, heads * d, device=DEV).to(torch.bfloat16)
    k = (torch.randn(frames * skv, heads * d, device=DEV) * 1.5).to(torch.bfloat16)
    v = torch.randn(frames * skv, heads * d, device=DEV).to(torch.float16)
    # make the chunks' maxima differ strongly: one dominant key late in the context for some rows
    k[skv - 5] *= 3.0
    out = torch.empty(frames * sq, heads * d, device=DEV, dtype=torch.bfloat16)
    chunks = skv // 512
    o_chunks = torch.empty(chunks, frames * sq, heads * d, device=DEV, dtype=torch.bfloat16)
    stats = torch.empty(chunks, frames * sq, heads, 2, device=DEV, dtype=torch.float32)
    _lib.call("rald_attn_d64_long", q.data_ptr(), heads * d, k.data_ptr(), heads * d, v.data_ptr(), heads * d,
              out.data_ptr(), heads * d, frames, heads, sq, skv, d ** -0.5, o_chunks.data_ptr(), stats.data_ptr(),
              _lib.cur_stream())
    qf = q.float().view(frames, sq, heads, d).transpose(1, 2)
    kf = k.float().view(frames, skv, heads, d).transpose(1, 2)
    vf = v.float().view(frames, skv, heads, d).transpose(1, 2)
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf).transpose(1, 2).reshape(frames * sq, -1)
    assert rel_l2(out, ref) < 1e-2


@torch.no_grad()
def test_noenc_denoiser_against_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "noenc.npz"))
    net = build_noenc(DEV)
    cube = torch.from_numpy(g["cube"]).to(DEV)
    tok = net.process_radar_cond(cube)
    assert tok.shape == (1, 2048, 512)
    assert rel_l2(tok[:, :64], torch.from_numpy(g["tokens_head"])) < 1e-5
    lat = synth.unit_latents([0]).to(DEV)
    for sigma in (80.0, 1.5):
        out = net(lat * sigma, torch.tensor(sigma), cube, "radar")
        assert rel_l2(out, torch.from_numpy(g[f"denoised_{sigma}"])) < 1e-2
    # two frames (batch path, per-frame context rows) reproduce the single frame
    out2 = net(torch.cat([lat, lat]) * 1.5, torch.tensor(1.5), torch.cat([cube, cube]), "radar")
    assert rel_l2(out2[1], torch.from_numpy(g["denoised_1.5"])[0]) < 1e-2 and torch.equal(out2[0], out2[1])
    x = net.sample(cube, batch_seeds=torch.tensor([0]), cond_type="radar")
    assert x.shape == (1, 512, 32) and torch.isfinite(x).all()
