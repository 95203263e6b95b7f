"""Host-side runtime of the radar-cube encoder: packs the fp32 parameters of ``models_radar_encoder.Encoder`` into
the layouts of csrc/conv3d.cu / csrc/encoder.cu (bf16, tap-major K), owns the workspaces and calls
``rald_radar_encoder`` (include/rald_b200.h). No hot-path arithmetic happens here."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import c_void_p

MAX_LEVELS, MAX_BLOCKS = 8, 4
i32, f32, i64 = ctypes.c_int32, ctypes.c_float, ctypes.c_int64


class EncConv(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("cin", i32), ("cout", i32), ("w_rows", i32), ("_pad", i32)]


class EncNorm(ctypes.Structure):
    _fields_ = [("g", c_void_p), ("b", c_void_p)]


class EncResBlock(ctypes.Structure):
    _fields_ = [("n1", EncNorm), ("c1", EncConv), ("n2", EncNorm), ("c2", EncConv), ("nin", EncConv)]


class EncAttnBlock(ctypes.Structure):
    _fields_ = [("n", EncNorm), ("qkv", EncConv), ("proj", EncConv)]


class EncLevel(ctypes.Structure):
    _fields_ = [("n_blocks", i32), ("n_attn", i32), ("block", EncResBlock * MAX_BLOCKS),
                ("attn", EncAttnBlock * MAX_BLOCKS), ("down", EncConv)]


class EncWeights(ctypes.Structure):
    _fields_ = [("n_levels", i32), ("in_ch", i32), ("ch", i32), ("z_ch", i32), ("groups", i32), ("_pad", i32),
                ("eps", f32), ("_pad2", i32), ("conv_in_w", c_void_p), ("conv_in_b", c_void_p),
                ("level", EncLevel * MAX_LEVELS), ("mid1", EncResBlock), ("mid2", EncResBlock),
                ("mid_attn", EncAttnBlock), ("norm_out", EncNorm), ("conv_out", EncConv)]


class EncWorkspace(ctypes.Structure):
    _fields_ = [("max_frames", i32), ("_pad", i32), ("elems", i64), ("x", c_void_p), ("y", c_void_p),
                ("t", c_void_p), ("xb", c_void_p), ("stats", c_void_p)]


def encoder_microbatch() -> int:
    return int(os.environ.get("RALD_B200_ENC_MICROBATCH", "32"))


def _n_tile(cout: int) -> int:
    return 128 if cout >= 128 else (64 if cout >= 64 else 32)


class EncoderRuntime(_lib.RuntimeNotCopied):
    def __init__(self, module):
        self.module = module
        self._sig = None
        self._keep = []   # packed tensors referenced by the ctypes struct
        self._ws = None
        self._ws_key = None

    def _signature(self):
        ps = list(self.module.parameters())
        return (ps[0].device, sum(p._version for p in ps), len(ps))

    # ------------------------------------------------------------------ packing
    def _conv3(self, conv) -> EncConv:
        w = conv.weight.detach()
        cout, cin = w.shape[0], w.shape[1]
        if tuple(w.shape[2:]) != (3, 3, 3):
            raise _lib.RaldError(f"expected a 3x3x3 convolution, got kernel {tuple(w.shape[2:])}")
        rows = -(-cout // _n_tile(cout)) * _n_tile(cout)
        wp = torch.zeros(rows, 27 * cin, device=w.device, dtype=torch.bfloat16)
        wp[:cout] = w.permute(0, 2, 3, 4, 1).reshape(cout, 27 * cin).to(torch.bfloat16)
        bp = torch.zeros(rows, device=w.device, dtype=torch.float32)
        bp[:cout] = conv.bias.detach().float()
        self._keep += [wp, bp]
        return EncConv(wp.data_ptr(), bp.data_ptr(), cin, cout, rows, 0)

    def _conv1(self, weights, biases) -> EncConv:
        """1x1x1 convolution(s) as a GEMM weight [sum(cout), cin] bf16 (q|k|v are concatenated)."""
        w = torch.cat([x.detach().reshape(x.shape[0], x.shape[1]) for x in weights]).to(torch.bfloat16).contiguous()
        b = torch.cat([x.detach().float() for x in biases]).contiguous()
        self._keep += [w, b]
        return EncConv(w.data_ptr(), b.data_ptr(), w.shape[1], w.shape[0], w.shape[0], 0)

    def _norm(self, gn) -> EncNorm:
        g, b = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
        self._keep += [g, b]
        return EncNorm(g.data_ptr(), b.data_ptr())

    def _resblock(self, rb) -> EncResBlock:
        if hasattr(rb, "conv_shortcut"):
            raise _lib.RaldError("ResnetBlock(conv_shortcut=True) is not used by any reference configuration")
        out = EncResBlock()
        out.n1, out.c1, out.n2, out.c2 = self._norm(rb.norm1), self._conv3(rb.conv1), self._norm(rb.norm2), \
            self._conv3(rb.conv2)
        if hasattr(rb, "nin_shortcut"):
            out.nin = self._conv1([rb.nin_shortcut.weight], [rb.nin_shortcut.bias])
        return out

    def _attnblock(self, ab) -> EncAttnBlock:
        out = EncAttnBlock()
        out.n = self._norm(ab.norm)
        out.qkv = self._conv1([ab.q.weight, ab.k.weight, ab.v.weight], [ab.q.bias, ab.k.bias, ab.v.bias])
        out.proj = self._conv1([ab.proj_out.weight], [ab.proj_out.bias])
        return out

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
        if m.num_resolutions > MAX_LEVELS or m.num_res_blocks > MAX_BLOCKS:
            raise _lib.RaldError("radar encoder deeper than the C ABI tables (8 levels x 4 blocks)")
        self._keep = []
        w = EncWeights()
        with torch.no_grad():
            w.n_levels, w.in_ch, w.ch, w.z_ch = m.num_resolutions, m.in_channels, m.ch, m.z_channels
            w.groups, w.eps = m.norm_out.num_groups, float(m.norm_out.eps)
            ci_w = m.conv_in.weight.detach().float().contiguous()
            ci_b = m.conv_in.bias.detach().float().contiguous()
            self._keep += [ci_w, ci_b]
            w.conv_in_w, w.conv_in_b = ci_w.data_ptr(), ci_b.data_ptr()
            for l, stage in enumerate(m.down):
                lv = w.level[l]
                lv.n_blocks, lv.n_attn = len(stage.block), len(stage.attn)
                if lv.n_attn not in (0, lv.n_blocks):
                    raise _lib.RaldError("radar encoder: attention must follow every block of a level or none")
                for j, rb in enumerate(stage.block):
                    lv.block[j] = self._resblock(rb)
                for j, ab in enumerate(stage.attn):
                    lv.attn[j] = self._attnblock(ab)
                if hasattr(stage, "downsample"):
                    if not stage.downsample.with_conv:
                        raise _lib.RaldError("Downsample(with_conv=False) is not used by any reference configuration")
                    lv.down = self._conv3(stage.downsample.conv)
            w.mid1, w.mid2 = self._resblock(m.mid.block_1), self._resblock(m.mid.block_2)
            w.mid_attn = self._attnblock(m.mid.attn_1)
            w.norm_out = self._norm(m.norm_out)
            w.conv_out = self._conv3(m.conv_out)
        self.weights = w
        self.device = dev
        self._sig = sig

    def _workspace(self, frames: int, vox: int) -> EncWorkspace:
        m = self.module
        mb = max(1, min(encoder_microbatch(), frames))
        # widest activation of any level, per frame (also covers the attention q|k|v buffer)
        need, v = vox * m.ch, vox
        for l, mult in enumerate(m.ch_mult):
            need = max(need, v * m.ch * mult)
            v //= 8
        v_last = vox // (8 ** (m.num_resolutions - 1))
        need = max(need, v_last * 3 * m.ch * m.ch_mult[-1])
        key = (mb, need)
        if self._ws_key != key:
            dev = self.device
            n = mb * need
            bufs = dict(x=torch.empty(n, device=dev, dtype=torch.float32),
                        y=torch.empty(n, device=dev, dtype=torch.float32),
                        t=torch.empty(n, device=dev, dtype=torch.float32),
                        xb=torch.empty(n, device=dev, dtype=torch.bfloat16),
                        stats=torch.empty(mb * self.weights.groups * 2, device=dev, dtype=torch.float64))
            ws = EncWorkspace()
            ws.max_frames, ws.elems = mb, n
            for k, t in bufs.items():
                setattr(ws, k, t.data_ptr())
            self._ws, self._ws_bufs, self._ws_key = ws, bufs, key
        return self._ws

    def forward(self, x_cl: torch.Tensor) -> torch.Tensor:
        """x_cl fp32 [B, D, H, W, Cin] on the device -> fp32 [B, D/s, H/s, W/s, z]."""
        self.ensure_packed()
        m = self.module
        B, D, H, W, cin = x_cl.shape
        if cin != m.in_channels:
            raise ValueError(f"radar encoder expects {m.in_channels} input channel(s), got {cin}")
        s = 2 ** (m.num_resolutions - 1)
        if D % s or H % s or W % s:
            raise ValueError(f"resolution {(D, H, W)} is not divisible by {s}")
        x_cl = x_cl.contiguous().float()
        out = torch.empty(B, D // s, H // s, W // s, m.z_channels, device=self.device, dtype=torch.float32)
        ws = self._workspace(B, D * H * W)
        _lib.call("rald_radar_encoder", ctypes.addressof(self.weights), ctypes.addressof(ws), x_cl.data_ptr(),
                  out.data_ptr(), B, D, H, W, _lib.cur_stream())
        return out


def _runtime(module) -> EncoderRuntime:
    rt = module.__dict__.get("_rt")
    if rt is None:
        rt = EncoderRuntime(module)
        module.__dict__["_rt"] = rt
    return rt


@torch.no_grad()
def encoder_forward(module, x: torch.Tensor, channels_last_in: bool = False, channels_last_out: bool = False):
    if x.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    if not channels_last_in:
        x = x.permute(0, 2, 3, 4, 1)
    out = _runtime(module).forward(x)
    return out if channels_last_out else out.permute(0, 4, 1, 2, 3).contiguous()
