"""Training forward / backward of the radar-cube encoder on the sm_100a kernels (``unfreeze_radar_enc: true``, the
default of the shipped configuration: the reference trains ``EDMPrecond.radar_enc`` jointly with the denoiser,
model/models_radar_generation.py:337-344, 378-388).

One ``torch.autograd.Function`` around ``Encoder.forward`` (model/models_radar_encoder.py:216-241):

* forward: the kernel sequence of csrc/encoder.cu issued from here, keeping what the backward pass needs (fp32 block
  inputs, GroupNorm statistics, the bf16 operands of every convolution, the attention blocks' q | k | v);
* backward, op by op in reverse:
    3x3x3 convolution  dgrad = ``rald_conv3d_cl`` on spatially flipped, in/out-transposed weights (stride 2: on the
                       output gradient zero-stuffed onto the input grid, ``rald_enc_stuff``); wgrad = ONE split-K GEMM
                       launch ``dW^T[kw, ci ; tap, co] = sum_P X_kw[ci][P] dY[co][P - off(tap)]`` over all voxels: the three
                       kw-shifted copies of X and dY are written once, transposed, onto the zero-padded voxel grid
                       (``rald_enc_pad_transpose``) where a (kd, kh) tap is a constant, 16-byte aligned index offset of
                       the dY^T boxes (``rald_gemm_bf16_accum_taps``, four taps per 256-wide tile);
                       bias gradient from the same pass over dY;
    GroupNorm(+swish)  ``rald_gn_bwd`` (two passes; adds the identity-shortcut gradient);
    1x1x1 convolutions GEMMs (dgrad against the transposed weight, wgrad over K = voxels with transposed operands);
    AttnBlock          ``rald_enc_attn_bwd`` between the GEMMs of its 1x1 convolutions.

bf16 tensor-core operands, fp32 accumulation, fp32 activations / gradients between the ops — as in the forward pass.
Host-side sequencing only; no torch fallback for any convolution, normalisation, attention or matrix product.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

# K-panel length of the conv weight-gradient operands (0: row-major operands with one row per channel)
_WGRAD_PANEL = int(os.environ.get("RALD_B200_WGRAD_PANEL", "32768"))

BF = torch.bfloat16
F32 = torch.float32


def _n_tile(cout: int) -> int:
    return 128 if cout >= 128 else (64 if cout >= 64 else 32)


def _s():
    return _lib.cur_stream()


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class _Conv3:
    """Packed forms of one 3x3x3 convolution: forward [rows][27*cin] (tap-major K) and dgrad [rows'][27*cout_pad]."""

    def __init__(self, name: str, conv, dev):
        w = conv.weight.detach().float()
        self.name = name
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.stride = conv.stride[0]
        self.w32 = w.contiguous()
        self.b32 = conv.bias.detach().float().contiguous()
        nt = _n_tile(self.cout)
        self.rows = -(-self.cout // nt) * nt
        wp = torch.zeros(self.rows, 27 * self.cin, device=dev, dtype=BF)
        wp[:self.cout] = w.permute(0, 2, 3, 4, 1).reshape(self.cout, 27 * self.cin).to(BF)
        bp = torch.zeros(self.rows, device=dev, dtype=F32)
        bp[:self.cout] = self.b32
        self.w_fwd, self.b_fwd = wp, bp
        if self.cin % 64 == 0:   # dgrad as a convolution: channels swap roles, taps are mirrored
            self.cout_pad = -(-self.cout // 64) * 64
            wd = torch.zeros(self.cin, self.cout_pad, 3, 3, 3, device=dev, dtype=F32)
            wd[:, :self.cout] = w.flip(2, 3, 4).transpose(0, 1)
            nt2 = _n_tile(self.cin)
            self.rows_d = -(-self.cin // nt2) * nt2
            wdp = torch.zeros(self.rows_d, 27 * self.cout_pad, device=dev, dtype=BF)
            wdp[:self.cin] = wd.permute(0, 2, 3, 4, 1).reshape(self.cin, 27 * self.cout_pad).to(BF)
            self.w_dgrad = wdp
            self.b_dgrad = torch.zeros(self.rows_d, device=dev, dtype=F32)


class _Lin:
    """1x1x1 convolution(s) as one GEMM weight [sum cout, cin] (+ its transpose for the dgrad)."""

    def __init__(self, names: List[str], convs, dev):
        self.names = names
        self.couts = [c.weight.shape[0] for c in convs]
        w = torch.cat([c.weight.detach().reshape(c.weight.shape[0], c.weight.shape[1]) for c in convs]).float()
        self.w = w.to(BF).contiguous()
        self.w_t = w.t().to(BF).contiguous()
        self.b = torch.cat([c.bias.detach().float() for c in convs]).contiguous()
        self.cout, self.cin = w.shape


class EncoderTrainRuntime(_lib.RuntimeNotCopied):
    def __init__(self, module):
        self.module = module
        self._sig = None

    def _signature(self):
        ps = list(self.module.parameters())
        return (ps[0].device, sum(p._version for p in ps), len(ps))

    # ------------------------------------------------------------------ packing
    def ensure_packed(self):
        sig = self._signature()
        if sig == self._sig:
            return
        m = self.module
        dev = sig[0]
        if dev.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only: move the module to a B200 (no CPU fallback)")
        if m.ch % 64 != 0:
            raise _lib.RaldError(f"radar encoder width ch={m.ch}: the tcgen05 convolution needs a multiple of 64")
        self.dev = dev
        self.groups, self.eps = m.norm_out.num_groups, float(m.norm_out.eps)
        with torch.no_grad():
            self.conv_in_w = m.conv_in.weight.detach().float().contiguous()
            self.conv_in_b = m.conv_in.bias.detach().float().contiguous()

            def res(prefix, rb):
                if hasattr(rb, "conv_shortcut"):
                    raise _lib.RaldError("ResnetBlock(conv_shortcut=True) is not used by any reference configuration")
                d = dict(prefix=prefix, n1=self._norm(prefix + "norm1", rb.norm1), c1=_Conv3(prefix + "conv1", rb.conv1, dev),
                         n2=self._norm(prefix + "norm2", rb.norm2), c2=_Conv3(prefix + "conv2", rb.conv2, dev), nin=None)
                if hasattr(rb, "nin_shortcut"):
                    d["nin"] = _Lin([prefix + "nin_shortcut"], [rb.nin_shortcut], dev)
                return d

            def attn(prefix, ab):
                return dict(prefix=prefix, n=self._norm(prefix + "norm", ab.norm),
                            qkv=_Lin([prefix + "q", prefix + "k", prefix + "v"], [ab.q, ab.k, ab.v], dev),
                            proj=_Lin([prefix + "proj_out"], [ab.proj_out], dev))
            self.levels = []
            for l, stage in enumerate(m.down):
                lv = dict(blocks=[res(f"down.{l}.block.{j}.", rb) for j, rb in enumerate(stage.block)],
                          attns=[attn(f"down.{l}.attn.{j}.", ab) for j, ab in enumerate(stage.attn)], down=None)
                if len(lv["attns"]) not in (0, len(lv["blocks"])):
                    raise _lib.RaldError("radar encoder: attention must follow every block of a level or none")
                if hasattr(stage, "downsample"):
                    if not stage.downsample.with_conv:
                        raise _lib.RaldError("Downsample(with_conv=False) is not used by any reference configuration")
                    lv["down"] = _Conv3(f"down.{l}.downsample.conv", stage.downsample.conv, dev)
                self.levels.append(lv)
            self.mid1, self.mid2 = res("mid.block_1.", m.mid.block_1), res("mid.block_2.", m.mid.block_2)
            self.mid_attn = attn("mid.attn_1.", m.mid.attn_1)
            self.norm_out = self._norm("norm_out", m.norm_out)
            self.conv_out = _Conv3("conv_out", m.conv_out, dev)
        self._sig = sig

    def _norm(self, name, gn):
        return dict(name=name, g=gn.weight.detach().float().contiguous(), b=gn.bias.detach().float().contiguous())

    # ------------------------------------------------------------------ primitives
    def _gn(self, x: torch.Tensor, nrm, mode: int):
        """x fp32 [B, V, C] -> (bf16 [B, V, C] = swish?(GN(x)), stats)."""
        B, V, C = x.shape
        stats = torch.empty(B, self.groups, 2, device=self.dev, dtype=torch.float64)
        _lib.call("rald_gn_stats", x.data_ptr(), B, V, C, self.groups, stats.data_ptr(), _s())
        out = torch.empty(B, V, C, device=self.dev, dtype=BF)
        _lib.call("rald_gn_apply", x.data_ptr(), stats.data_ptr(), nrm["g"].data_ptr(), nrm["b"].data_ptr(), out.data_ptr(),
                  B, V, C, self.groups, self.eps, mode, _s())
        return out, stats

    def _cast(self, x: torch.Tensor) -> torch.Tensor:
        B, V, C = x.shape
        out = torch.empty(B, V, C, device=self.dev, dtype=BF)
        _lib.call("rald_gn_apply", x.data_ptr(), 0, 0, 0, out.data_ptr(), B, V, C, 0, 0.0, 2, _s())
        return out

    def _conv3(self, xb: torch.Tensor, cv: _Conv3, dims, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = xb.shape[0]
        D, H, W = dims
        s = cv.stride
        out = torch.empty(B, (D // s) * (H // s) * (W // s), cv.cout, device=self.dev, dtype=F32)
        _lib.call("rald_conv3d_cl", xb.data_ptr(), cv.w_fwd.data_ptr(), cv.rows, cv.b_fwd.data_ptr(), _p(resid),
                  out.data_ptr(), B, D, H, W, cv.cin, cv.cout, s, _s())
        return out

    def _gn_bwd(self, x, stats, nrm, dy, swish: int, add: Optional[torch.Tensor], grads):
        B, V, C = x.shape
        sums = torch.empty(B, C, 2, device=self.dev, dtype=torch.float64)
        dx = torch.empty(B, V, C, device=self.dev, dtype=F32)
        _lib.call("rald_gn_bwd", x.data_ptr(), dy.data_ptr(), stats.data_ptr(), nrm["g"].data_ptr(), nrm["b"].data_ptr(), B, V,
                  C, self.groups, self.eps, swish, sums.data_ptr(), _p(add), dx.data_ptr(), _s())
        tot = sums.sum(0)                      # [C, 2]: B x C numbers
        grads[nrm["name"] + ".weight"] = tot[:, 0].float()
        grads[nrm["name"] + ".bias"] = tot[:, 1].float()
        return dx

    def _colsum(self, t2d: torch.Tensor) -> torch.Tensor:
        R, C = t2d.shape
        chunks = min(512, (R + 255) // 256)
        ws = torch.empty(chunks * C, device=self.dev, dtype=F32)
        out = torch.empty(C, device=self.dev, dtype=F32)
        _lib.call("rald_colsum", t2d.data_ptr(), 1 if t2d.dtype == F32 else 0, C, R, C, ws.data_ptr(), ws.numel(),
                  out.data_ptr(), 0, _s())
        return out

    def _transpose(self, t2d: torch.Tensor, want_plain: bool = False):
        R, C = t2d.shape
        is_f32 = t2d.dtype == F32
        plain = torch.empty(R, C, device=self.dev, dtype=BF) if (want_plain and is_f32) else None
        Rp = (R + 7) // 8 * 8
        tt = torch.empty(C, Rp, device=self.dev, dtype=BF) if Rp == R else torch.zeros(C, Rp, device=self.dev, dtype=BF)
        _lib.call("rald_cast_transpose", t2d.data_ptr(), 1 if is_f32 else 0, C, R, C, _p(plain), C, tt.data_ptr(), Rp, 0, _s())
        return (plain if is_f32 else (t2d if want_plain else None)), tt

    def _conv3_dgrad(self, dy16: torch.Tensor, cv: _Conv3, dims) -> torch.Tensor:
        """dy16 bf16 [B, V, cout_pad] on the INPUT grid `dims` -> d loss / d (conv input) fp32 [B, V, cin]."""
        B = dy16.shape[0]
        D, H, W = dims
        out = torch.empty(B, D * H * W, cv.cin, device=self.dev, dtype=F32)
        _lib.call("rald_conv3d_cl", dy16.data_ptr(), cv.w_dgrad.data_ptr(), cv.rows_d, cv.b_dgrad.data_ptr(), 0,
                  out.data_ptr(), B, D, H, W, cv.cout_pad, cv.cin, 1, _s())
        return out

    def _dy16(self, dy: torch.Tensor, cv: _Conv3) -> torch.Tensor:
        """fp32 [B, V, cout] -> bf16 [B, V, cout_pad] (zero channels appended when cout is not a multiple of 64)."""
        if cv.cout_pad == cv.cout:
            return self._cast(dy)
        B, V, _ = dy.shape
        out = torch.zeros(B, V, cv.cout_pad, device=self.dev, dtype=BF)
        out[:, :, :cv.cout] = dy.to(BF)          # conv_out only: 64 voxels x 16 channels per frame
        return out

    def _conv3_wgrad(self, dy: torch.Tensor, x_in: torch.Tensor, cv_cout: int, cin: int, dims, stride: int):
        """dy fp32 [B, Vo, cout] (output grid), x_in [B, V, cin] (bf16 or fp32, input grid `dims`) ->
        (dW fp32 [cout, cin, 3, 3, 3], db fp32 [cout])."""
        B = dy.shape[0]
        D, H, W = dims
        Wp = -(-(W + 2) // 8) * 8                       # padded row pitch: kd / kh tap offsets become multiples of 8
        padded = B * (D + 2) * (H + 2) * Wp
        cin_rows = -(-cin // 32) * 32
        co_rows = -(-cv_cout // 64) * 64                 # a tap = one box of co_rows rows of dY^T
        o = 1 if stride == 1 else 0
        per_tile = max(1, 256 // co_rows)
        n_taps = -(-9 // per_tile) * per_tile
        S1, S2 = (H + 2) * Wp, Wp
        offs = [-((kd - o) * S1 + (kh - o) * S2) for kd in range(3) for kh in range(3)] + [0] * (n_taps - 9)
        bsum = torch.empty(cv_cout, device=self.dev, dtype=torch.float64)      # bias gradient, from the same pass over dY
        shifts = (ctypes.c_int * n_taps)(*offs)
        dWt = torch.zeros(3 * cin_rows, n_taps * co_rows, device=self.dev, dtype=F32)
        if _WGRAD_PANEL > 0:
            PL = _WGRAD_PANEL                           # K-panel length: a tile's rows lie within a few 2 MB pages
            n_pan = -(-(padded + 8) // PL)
            K = n_pan * PL
            halo = -(-max(abs(v) for v in offs) // 8) * 8
            # panel-major operands: X copies [n_pan][3*cin_rows][PL], dY^T [n_pan][co_rows][halo | PL | halo]
            dyT = torch.zeros(n_pan, co_rows, PL + 2 * halo, device=self.dev, dtype=BF)
            xT = torch.zeros(n_pan, 3 * cin_rows, PL, device=self.dev, dtype=BF)    # [kw][cin_rows]: X pre-shifted by kw - o
            ld = 0
        else:
            PL = halo = n_pan = 0
            ld = K = -(-(padded + 8) // 16384) * 16384
            dyT = torch.zeros(co_rows, ld, device=self.dev, dtype=BF)
            xT = torch.zeros(3 * cin_rows, ld, device=self.dev, dtype=BF)
        _lib.call("rald_enc_pad_transpose", dy.data_ptr(), 1, B, D // stride, H // stride, W // stride, cv_cout, stride, Wp,
                  1, co_rows, 0, dyT.data_ptr(), ld, PL, halo, n_pan, bsum.data_ptr(), _s())
        _lib.call("rald_enc_pad_transpose", x_in.data_ptr(), 1 if x_in.dtype == F32 else 0, B, D, H, W, cin, 1, Wp, 3,
                  cin_rows, o, xT.data_ptr(), ld, PL, 0, n_pan, 0, _s())
        # dW^T in ONE launch: out[kw*cin_rows + ci][tap*co_rows + co] = sum_P X_kw[ci][P] dY[co][P - off(tap)], the nine
        # (kd, kh) taps as shifted boxes of dY^T stacked four (co_rows = 64) / two (128) to a 256-wide tile; the tap count is
        # padded to a multiple of that with dummy taps whose output is dropped
        _lib.call("rald_gemm_bf16_accum_taps", xT.data_ptr(), ld, dyT.data_ptr(), ld, co_rows, n_taps,
                  ctypes.addressof(shifts), PL, halo, dWt.data_ptr(), n_taps * co_rows, 3 * cin_rows, K, _s())
        # [kw][cin_rows][tap][co_rows] -> [cout][cin][kd][kh][kw]
        gw = dWt.reshape(3, cin_rows, n_taps, co_rows)[:, :cin, :9, :cv_cout].reshape(3, cin, 3, 3, cv_cout) \
            .permute(4, 1, 2, 3, 0).contiguous()
        return gw, bsum.float()

    def _lin_fwd(self, a16: torch.Tensor, lin: _Lin, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
        """a16 bf16 [R, cin] -> fp32 [R, cout] = a W^T + b (+ resid)."""
        R = a16.shape[0]
        out = torch.empty(R, lin.cout, device=self.dev, dtype=F32)
        _lib.call("rald_gemm_bf16", a16.data_ptr(), lin.cin, lin.w.data_ptr(), lin.cin, out.data_ptr(), lin.cout,
                  lin.b.data_ptr(), _p(resid), lin.cout if resid is not None else 0, R, lin.cout, lin.cin, 1, 0, _s())
        return out

    def _lin_bwd(self, dy: torch.Tensor, a16: torch.Tensor, lin: _Lin, grads, add: Optional[torch.Tensor] = None):
        """dy fp32 [R, cout], a16 bf16 [R, cin] (the forward input) -> d input fp32 [R, cin] (+ add); fills the weight and
        bias gradients of the (possibly concatenated) 1x1x1 convolutions."""
        R = dy.shape[0]
        dy16, dyT = self._transpose(dy, want_plain=True)
        _, aT = self._transpose(a16)
        gW = torch.zeros(lin.cout, lin.cin, device=self.dev, dtype=F32)
        _lib.call("rald_gemm_bf16_accum", dyT.data_ptr(), dyT.shape[1], aT.data_ptr(), aT.shape[1], gW.data_ptr(), lin.cin,
                  lin.cout, lin.cin, dyT.shape[1], _s())
        gb = self._colsum(dy)
        r0 = 0
        for name, co in zip(lin.names, lin.couts):
            grads[name + ".weight"] = gW[r0:r0 + co].reshape(co, lin.cin, 1, 1, 1)
            grads[name + ".bias"] = gb[r0:r0 + co]
            r0 += co
        dx = torch.empty(R, lin.cin, device=self.dev, dtype=F32)
        _lib.call("rald_gemm_bf16", dy16.data_ptr(), lin.cout, lin.w_t.data_ptr(), lin.cout, dx.data_ptr(), lin.cin, 0,
                  _p(add), lin.cin if add is not None else 0, R, lin.cin, lin.cout, 1, 0, _s())
        return dx

    # ------------------------------------------------------------------ blocks
    def _res_fwd(self, x, rb, dims, tape):
        B, V, _ = x.shape
        xb1, st1 = self._gn(x, rb["n1"], 0)
        t = self._conv3(xb1, rb["c1"], dims)
        xb2, st2 = self._gn(t, rb["n2"], 0)
        xc = None
        if rb["nin"] is not None:
            xc = self._cast(x)
            sc = self._lin_fwd(xc.reshape(B * V, -1), rb["nin"]).reshape(B, V, -1)
            out = self._conv3(xb2, rb["c2"], dims, resid=sc)
        else:
            out = self._conv3(xb2, rb["c2"], dims, resid=x)
        tape.append(("res", rb, dims, dict(x=x, xb1=xb1, st1=st1, t=t, xb2=xb2, st2=st2, xc=xc)))
        return out

    def _res_bwd(self, rb, dims, k, dout, grads):
        B, V, cout = dout.shape
        c1, c2 = rb["c1"], rb["c2"]
        gw, gb = self._conv3_wgrad(dout, k["xb2"], c2.cout, c2.cin, dims, 1)
        grads[c2.name + ".weight"], grads[c2.name + ".bias"] = gw, gb
        d_act2 = self._conv3_dgrad(self._dy16(dout, c2), c2, dims)
        dt = self._gn_bwd(k["t"], k["st2"], rb["n2"], d_act2, 1, None, grads)
        gw, gb = self._conv3_wgrad(dt, k["xb1"], c1.cout, c1.cin, dims, 1)
        grads[c1.name + ".weight"], grads[c1.name + ".bias"] = gw, gb
        d_act1 = self._conv3_dgrad(self._dy16(dt, c1), c1, dims)
        if rb["nin"] is not None:
            d_sc = self._lin_bwd(dout.reshape(B * V, cout), k["xc"].reshape(B * V, -1), rb["nin"], grads).reshape(B, V, -1)
            return self._gn_bwd(k["x"], k["st1"], rb["n1"], d_act1, 1, d_sc, grads)
        return self._gn_bwd(k["x"], k["st1"], rb["n1"], d_act1, 1, dout, grads)

    def _attn_fwd(self, x, ab, tape):
        B, n, C = x.shape
        if n > 64:
            raise _lib.RaldError(f"radar encoder: AttnBlock over {n} voxels (the kernel handles <= 64 = 8x4x2)")
        xb, st = self._gn(x, ab["n"], 1)
        qkv = self._lin_fwd(xb.reshape(B * n, C), ab["qkv"])
        a = torch.empty(B * n, C, device=self.dev, dtype=BF)
        _lib.call("rald_enc_attn", qkv.data_ptr(), a.data_ptr(), B, n, C, _s())
        out = self._lin_fwd(a, ab["proj"], resid=x.reshape(B * n, C)).reshape(B, n, C)
        tape.append(("attn", ab, None, dict(x=x, xb=xb, st=st, qkv=qkv, a=a)))
        return out

    def _attn_bwd(self, ab, k, dout, grads):
        B, n, C = dout.shape
        d_a = self._lin_bwd(dout.reshape(B * n, C), k["a"], ab["proj"], grads)
        dqkv = torch.empty(B * n, 3 * C, device=self.dev, dtype=F32)
        _lib.call("rald_enc_attn_bwd", k["qkv"].data_ptr(), d_a.data_ptr(), dqkv.data_ptr(), B, n, C, _s())
        d_xb = self._lin_bwd(dqkv, k["xb"].reshape(B * n, C), ab["qkv"], grads).reshape(B, n, C)
        return self._gn_bwd(k["x"], k["st"], ab["n"], d_xb, 0, dout, grads)

    # ------------------------------------------------------------------ forward / backward of the whole encoder
    def forward(self, x_cl: torch.Tensor):
        """x_cl fp32 [B, D, H, W, Cin] -> (fp32 [B, D/s, H/s, W/s, z], tape)."""
        self.ensure_packed()
        m = self.module
        B, D, H, W, cin = x_cl.shape
        if cin != m.in_channels:
            raise ValueError(f"radar encoder expects {m.in_channels} input channel(s), got {cin}")
        s = 2 ** (m.num_resolutions - 1)
        if D % s or H % s or W % s:
            raise ValueError(f"resolution {(D, H, W)} is not divisible by {s}")
        x_cl = x_cl.contiguous().float()
        tape: list = []
        x = torch.empty(B, D * H * W, m.ch, device=self.dev, dtype=F32)
        _lib.call("rald_enc_conv_in", x_cl.data_ptr(), self.conv_in_w.data_ptr(), self.conv_in_b.data_ptr(), x.data_ptr(), B, D,
                  H, W, cin, m.ch, _s())
        tape.append(("conv_in", None, (D, H, W), dict(x=x_cl)))
        dims = (D, H, W)
        for lv in self.levels:
            for j, rb in enumerate(lv["blocks"]):
                x = self._res_fwd(x, rb, dims, tape)
                if j < len(lv["attns"]):
                    x = self._attn_fwd(x, lv["attns"][j], tape)
            if lv["down"] is not None:
                xc = self._cast(x)
                x = self._conv3(xc, lv["down"], dims)
                tape.append(("down", lv["down"], dims, dict(xc=xc)))
                dims = tuple(d // 2 for d in dims)
        x = self._res_fwd(x, self.mid1, dims, tape)
        x = self._attn_fwd(x, self.mid_attn, tape)
        x = self._res_fwd(x, self.mid2, dims, tape)
        xb, st = self._gn(x, self.norm_out, 0)
        out = self._conv3(xb, self.conv_out, dims)
        tape.append(("out", None, dims, dict(x=x, xb=xb, st=st)))
        return out.reshape(B, dims[0], dims[1], dims[2], m.z_channels), tape

    def backward(self, tape, dfeat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """dfeat fp32 [B, d, h, w, z] -> {parameter name (relative to the Encoder): fp32 gradient}."""
        self.ensure_packed()
        grads: Dict[str, torch.Tensor] = {}
        B = dfeat.shape[0]
        dy = dfeat.reshape(B, -1, dfeat.shape[-1]).contiguous().float()
        for kind, obj, dims, k in reversed(tape):
            if kind == "out":
                cv = self.conv_out
                gw, gb = self._conv3_wgrad(dy, k["xb"], cv.cout, cv.cin, dims, 1)
                grads["conv_out.weight"], grads["conv_out.bias"] = gw, gb
                d_act = self._conv3_dgrad(self._dy16(dy, cv), cv, dims)
                dy = self._gn_bwd(k["x"], k["st"], self.norm_out, d_act, 1, None, grads)
            elif kind == "res":
                dy = self._res_bwd(obj, dims, k, dy, grads)
            elif kind == "attn":
                dy = self._attn_bwd(obj, k, dy, grads)
            elif kind == "down":
                cv = obj
                D, H, W = dims
                gw, gb = self._conv3_wgrad(dy, k["xc"], cv.cout, cv.cin, dims, 2)
                grads[cv.name + ".weight"], grads[cv.name + ".bias"] = gw, gb
                z = torch.zeros(B, D * H * W, cv.cout_pad, device=self.dev, dtype=BF)
                _lib.call("rald_enc_stuff", dy.data_ptr(), B, D // 2, H // 2, W // 2, cv.cout, z.data_ptr(), _s())
                dy = self._conv3_dgrad(z, cv, dims)
            elif kind == "conv_in":
                m = self.module
                gw, gb = self._conv3_wgrad(dy, k["x"].reshape(B, -1, m.in_channels), m.ch, m.in_channels, dims, 1)
                grads["conv_in.weight"], grads["conv_in.bias"] = gw, gb
        return grads


class EncoderTrainFunction(torch.autograd.Function):
    """feat = Encoder(x) (channels last in and out) with gradients for every encoder parameter."""

    @staticmethod
    def forward(ctx, rt: EncoderTrainRuntime, names: List[str], x_cl, *params):
        with torch.no_grad():
            out, tape = rt.forward(x_cl)
        ctx.rt, ctx.names, ctx.tape = rt, names, tape
        return out

    @staticmethod
    def backward(ctx, dfeat):
        with torch.no_grad():
            grads = ctx.rt.backward(ctx.tape, dfeat)
        ctx.tape = None
        return (None, None, None, *[grads[n] for n in ctx.names])
