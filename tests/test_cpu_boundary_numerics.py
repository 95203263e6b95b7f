"""CPU: the NUMERICAL DESIGN of the evaluation-boundary kernel (csrc/dit_misc.cu), restated in numpy and held against
fp64: split-bf16 operands (hi = truncation, lo = round-to-nearest of the exact remainder; weights hi = round-to-nearest),
the three products hi*hi + hi*lo + lo*hi, the per-row shift K (median of three samples of the row), the one-pass
shifted moments and the algebraic LayerNorm  F = rstd (acc - mean' colsum(W')) + ln_b W_out^T.
This pins the arithmetic the CUDA kernel implements (tests/test_gpu_boundary.py pins the kernel itself against fp64 on
the GPU): the 3e-5 bar of that test must hold for well-behaved rows AND for the adversarial ones — a row mean tens of
standard deviations from zero, a massive channel among the samples the shift is taken from, constant rows. (Written
before the kernel took the median: with K = the mean of the row's first four values this file's massive-channel case
failed at 7e-5 — the split operands carry 16+ bits relative to |h - K|, not to |h - mean|.)"""
import numpy as np
import pytest


K_COLS = (5, 173, 347)   # BD_K0 / BD_K1 / BD_K2 of dit_misc.cu


def _bf16_trunc(x):
    b = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFF0000)
    return b.view(np.float32)


def _bf16_rn(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _split_act(x):          # split_pair() of ptx.cuh
    hi = _bf16_trunc(x)
    lo = _bf16_rn((x - hi).astype(np.float32))
    return hi, lo


def _split_w(w):            # boundary_pack_kernel
    hi = _bf16_rn(w)
    lo = _bf16_rn((w - hi).astype(np.float32))
    return hi, lo


def kernel_restatement(h, ln_w, ln_b, w_out):
    """h [T, 512] fp32, w_out [C, 512] fp32 -> F [T, C] as the kernel computes it (products exact, accumulation in
    fp64 rounded once: the tensor core's fp32 accumulation adds ~1e-7 on top)."""
    h = h.astype(np.float32)
    wp = (w_out.T.astype(np.float32) * ln_w.astype(np.float32)[:, None]).astype(np.float32)     # W' [512, C]
    cs = wp.sum(0, dtype=np.float32)
    bw = (ln_b.astype(np.float32)[:, None] * w_out.T.astype(np.float32)).sum(0, dtype=np.float32)
    K = np.median(h[:, list(K_COLS)], axis=1).astype(np.float32)        # median3() of dit_misc.cu
    a = (h - K[:, None]).astype(np.float32)
    ahi, alo = _split_act(a)
    whi, wlo = _split_w(wp)
    acc = (alo.astype(np.float64) @ whi.astype(np.float64) + ahi.astype(np.float64) @ wlo.astype(np.float64)
           + ahi.astype(np.float64) @ whi.astype(np.float64)).astype(np.float32)
    s1 = a.sum(1, dtype=np.float32)
    s2 = (a * a).sum(1, dtype=np.float32)
    m = s1 * np.float32(1 / 512)
    var = np.maximum(s2 * np.float32(1 / 512) - m * m, np.float32(0))
    rstd = (1.0 / np.sqrt(var + np.float32(1e-5))).astype(np.float32)
    return (rstd[:, None] * (acc - m[:, None] * cs[None, :]) + bw[None, :]).astype(np.float32)


def reference(h, ln_w, ln_b, w_out):
    h = h.astype(np.float64)
    mu = h.mean(1, keepdims=True)
    var = ((h - mu) ** 2).mean(1, keepdims=True)
    y = (h - mu) / np.sqrt(var + 1e-5) * ln_w.astype(np.float64) + ln_b.astype(np.float64)
    return y @ w_out.T.astype(np.float64)


def _weights(rng, C=32):
    return (1 + 0.2 * rng.standard_normal(512)).astype(np.float32), (0.1 * rng.standard_normal(512)).astype(np.float32), \
        (rng.standard_normal((C, 512)) / 512 ** 0.5).astype(np.float32)


def _err(F, ref):
    return float(np.abs(F.astype(np.float64) - ref).max() / np.abs(ref).max())


def test_bf16_emulation_is_exact():
    x = np.array([1.0, 1.00390625, -3.14159, 1e-20, 65504.0, 0.1], np.float32)
    hi, lo = _split_act(x)
    assert np.all((hi.view(np.uint32) & 0xFFFF) == 0) and np.all((lo.view(np.uint32) & 0xFFFF) == 0)
    assert np.all(np.abs(x - (hi + lo)) <= np.abs(x) * 2.0 ** -16)       # two halves carry 16+ mantissa bits
    assert _bf16_rn(np.float32(1.00390625)) == np.float32(1.0)            # tie -> even
    assert _bf16_rn(np.float32(1.01171875)) == np.float32(1.015625)       # tie -> even (upwards)


@pytest.mark.parametrize("case", ["plain", "large_mean", "massive_channel_in_shift", "massive_channel_elsewhere",
                                  "tiny_scale", "huge_scale"])
def test_boundary_arithmetic_meets_its_bar(case):
    rng = np.random.default_rng(sum(ord(c) for c in case))
    T = 256
    ln_w, ln_b, w_out = _weights(rng)
    h = rng.standard_normal((T, 512)) * 3.0
    if case == "large_mean":
        h += 40.0 * rng.standard_normal((T, 1)) + 100.0      # |mean| = 30 - 50 standard deviations
    elif case == "massive_channel_in_shift":
        h[:, K_COLS[0]] += 400.0                              # one of the three samples is 130 sigma out: the median ignores it
    elif case == "massive_channel_elsewhere":
        h[:, 300] -= 900.0
    elif case == "tiny_scale":
        h *= 1e-3
    elif case == "huge_scale":
        h *= 1e4
    F = kernel_restatement(h.astype(np.float32), ln_w, ln_b, w_out)
    ref = reference(h.astype(np.float32), ln_w, ln_b, w_out)
    e = _err(F, ref)
    print(case, f"{e:.2e}")
    assert e < 3e-5


def test_constant_rows_do_not_blow_up():
    """var = 0: the clamp keeps rstd = 1 / sqrt(eps) finite and F = ln_b W_out^T as LayerNorm gives."""
    rng = np.random.default_rng(7)
    ln_w, ln_b, w_out = _weights(rng)
    h = np.full((16, 512), 3.25, np.float32)
    F = kernel_restatement(h, ln_w, ln_b, w_out)
    ref = reference(h, ln_w, ln_b, w_out)
    assert np.isfinite(F).all() and _err(F, ref) < 1e-5


def test_conv_in_split_products_meet_their_bar():
    """conv_in_mma_kernel (csrc/enc_misc.cu): 27 taps of a cube value in [0, 1] times fp32 weights as split-bf16 products
    (activations hi = truncation, weights hi = round-to-nearest), against fp64 — the 2e-5 rel-L2 bar of
    tests/test_gpu_encoder.py::test_conv_in_against_torch with an order of magnitude to spare."""
    rng = np.random.default_rng(11)
    x = rng.random((4096, 27)).astype(np.float32)                 # im2col rows: one output voxel's 27 taps
    w = (rng.standard_normal((27, 64)) * 0.2).astype(np.float32)
    b = rng.standard_normal(64).astype(np.float32)
    xhi, xlo = _split_act(x)
    whi, wlo = _split_w(w)
    out = (xlo.astype(np.float64) @ whi.astype(np.float64) + xhi.astype(np.float64) @ wlo.astype(np.float64)
           + xhi.astype(np.float64) @ whi.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
    ref = x.astype(np.float64) @ w.astype(np.float64) + b.astype(np.float64)
    e = float(np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref))
    print("conv_in split-bf16 rel-L2", f"{e:.2e}")
    assert e < 5e-6
