"""GPU parity of the training / AE-evaluation loop helpers through the C ABI: the fused
multi-tensor EMA update and the occupancy accuracy / IoU kernel, bit-exact against the fixture written from the
unmodified reference code and against the oracle on seeded inputs (ragged lists, misaligned views, empty tensors, a
full-size denoiser parameter list)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import rald_oracle as orc
from rald_b200 import _lib, postproc, train

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _g():
    return np.load(os.path.join(GOLDEN, "trainloop.npz"))


def test_update_ema_fixture_bit_exact():
    g = _g()
    n = sum(1 for k in g.files if k.startswith("ema_target_"))
    src = [torch.from_numpy(g[f"ema_source_{i}"]).to(DEV) for i in range(n)]
    t = [torch.from_numpy(g[f"ema_target_{i}"]).to(DEV) for i in range(n)]
    before = _lib.launch_count()
    for _ in range(3):
        train.update_ema(t, src, rate=float(g["ema_rate"]))
    assert _lib.launch_count() - before == 3                       # one launch per update for the whole list
    for i in range(n):
        assert np.array_equal(t[i].cpu().numpy(), g[f"ema_after3_{i}"]), i
    t = [torch.from_numpy(g[f"ema_target_{i}"]).to(DEV) for i in range(n)]
    train.update_ema(t, src)                                       # default rate 0.99
    for i in range(n):
        assert np.array_equal(t[i].cpu().numpy(), g[f"ema_default_{i}"]), i


def test_update_ema_ragged_misaligned_and_empty():
    gen = torch.Generator().manual_seed(3)
    sizes = [0, 1, 4095, 4096, 4097, 3 * 4096 + 5, 0, 70001]
    base_t = [torch.randn(s + 3, generator=gen) for s in sizes]
    base_s = [torch.randn(s + 3, generator=gen) for s in sizes]
    # odd element offsets: pointers that are 4- but not 16-byte aligned take the scalar path
    off = [0, 1, 2, 3, 1, 0, 2, 3]
    t_cpu = [b[o:o + s].clone() for b, o, s in zip(base_t, off, sizes)]
    s_cpu = [b[o:o + s].clone() for b, o, s in zip(base_s, off, sizes)]
    dev_t = [b.to(DEV) for b in base_t]
    dev_s = [b.to(DEV) for b in base_s]
    t_dev = [b[o:o + s] for b, o, s in zip(dev_t, off, sizes)]
    s_dev = [b[o:o + s] for b, o, s in zip(dev_s, off, sizes)]
    want = [t.numpy().copy() for t in t_cpu]
    for _ in range(2):
        orc.update_ema(want, [s.numpy() for s in s_cpu], rate=0.999)
        train.update_ema(t_dev, s_dev, rate=0.999)
    for i, (a, b) in enumerate(zip(t_dev, want)):
        assert np.array_equal(a.cpu().numpy(), b), i
    # elements outside the views are untouched
    for b_dev, b_cpu, o, s in zip(dev_t, base_t, off, sizes):
        assert torch.equal(b_dev[:o].cpu(), b_cpu[:o]) and torch.equal(b_dev[o + s:].cpu(), b_cpu[o + s:])


def test_update_ema_full_parameter_list_matches_torch():
    """The default denoiser's 637-tensor list: one launch, equal to torch's own per-tensor CUDA kernels bit for bit."""
    from helpers import build_denoiser
    net = build_denoiser(device=DEV)
    params = [p for p in net.parameters()]
    gen = torch.Generator(device=DEV).manual_seed(1)
    ema = [torch.randn(p.shape, device=DEV, generator=gen) for p in params]
    ref = [e.clone() for e in ema]
    for targ, src in zip(ref, params):
        targ.detach().mul_(0.9999).add_(src.detach(), alpha=1 - 0.9999)
    before = _lib.launch_count()
    train.update_ema(ema, params, rate=0.9999)
    assert _lib.launch_count() - before == 1
    assert all(torch.equal(a, b) for a, b in zip(ema, ref))
    with pytest.raises(_lib.RaldError):
        train.update_ema([ema[0].double()], [params[0]])
    with pytest.raises(_lib.RaldError):
        train.update_ema([ema[0]], [params[1]])


def test_occupancy_iou_fixture_and_seeded():
    g = _g()
    acc, iou = postproc.occupancy_iou(torch.from_numpy(g["iou_logits"]).to(DEV),
                                      torch.from_numpy(g["iou_labels"]).to(DEV), 0.0)
    assert np.array_equal(acc.cpu().numpy(), g["iou_accuracy"])
    assert np.array_equal(iou.cpu().numpy(), g["iou_iou"], equal_nan=True)
    gen = torch.Generator().manual_seed(9)
    for B, Q, thr in ((1, 1, 0.0), (3, 4096, 0.1), (7, 500000, 0.0)):
        lg = torch.randn(B, Q, generator=gen)
        lb = (torch.rand(B, Q, generator=gen) < 0.1).float()
        a0, i0 = orc.occupancy_iou(lg, lb, thr)
        a1, i1 = postproc.occupancy_iou(lg.to(DEV).unsqueeze(-1), lb.to(DEV), thr)
        assert np.array_equal(a1.cpu().numpy(), a0.numpy()) and np.array_equal(i1.cpu().numpy(), i0.numpy(), equal_nan=True)


def test_trainloop_abi_argument_checks():
    lib = _lib.lib()
    assert lib.rald_ema_update(0, 1, 1, 1, 0.5, 0.5, 0) != 0 and b"null" in lib.rald_last_error()
    t = torch.zeros(16, device=DEV, dtype=torch.int64)
    assert lib.rald_ema_update(t.data_ptr(), 0, 1, 1, 0.5, 0.5, 0) != 0
    assert lib.rald_occupancy_iou(0, 0, 1, 8, 0.0, 0, 0, 0) != 0
    f = torch.zeros(8, device=DEV)
    assert lib.rald_occupancy_iou(f.data_ptr(), f.data_ptr(), 0, 8, 0.0, f.data_ptr(), t.data_ptr(), 0) != 0


def test_update_ema_into_module_parameters_repacks_runtime():
    """the fused EMA writes through raw pointers; the runtimes notice weight changes through autograd
    version counters, so update_ema must bump them (ema_model = deepcopy(model) pattern)."""
    import copy
    from helpers import build_denoiser
    from rald_b200 import synth
    net = build_denoiser(name="kl_d512_m512_l32_d12_edm", device="cuda")
    ema = copy.deepcopy(net)
    tok = synth.unit_latents([7])[:, :64, :].repeat(1, 1, 16).cuda()      # [1, 64, 512] conditioning tokens
    x = synth.unit_latents([0]).cuda()
    sigma = torch.tensor(1.5)
    with torch.no_grad():
        before = ema(x * 1.5, sigma, tok, "radar").clone()
        for p in net.parameters():
            p.mul_(1.05)
        train.update_ema(list(ema.parameters()), list(net.parameters()), rate=0.5)
        after = ema(x * 1.5, sigma, tok, "radar")
        fresh = copy.deepcopy(ema)              # same weights, freshly packed
        expect = fresh(x * 1.5, sigma, tok, "radar")
    assert not torch.equal(before, after)
    assert torch.equal(after, expect)
