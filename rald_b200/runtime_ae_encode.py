"""Host-side runtime of KLAutoEncoder.encode: packs the encoder-side weights, owns the per-frame workspace and
calls ``rald_ae_encode_stats`` / ``rald_ae_posterior`` (include/rald_b200.h). No hot-path arithmetic happens here."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_void_p
from .runtime_dit import geglu_pack_index

i32, f32 = ctypes.c_int32, ctypes.c_float
QUERY_TYPES = {"point": 0, "learnable": 1, "mix": 2}


class AeEncWeights(ctypes.Structure):
    _fields_ = [("dim", i32), ("n_latents", i32), ("latent_dim", i32), ("heads", i32), ("query_type", i32),
                ("stats_rows", i32), ("freq24", f32 * 24),
                ("wpe", c_void_p), ("pe_bias", c_void_p), ("mix_q", c_void_p), ("mix_wkv", c_void_p),
                ("mix_wo", c_void_p), ("mix_bo", c_void_p), ("s_latents", c_void_p), ("wproj", c_void_p),
                ("bproj", c_void_p), ("latents", c_void_p),
                ("ca_ln_w", c_void_p), ("ca_ln_b", c_void_p), ("ca_lnc_w", c_void_p), ("ca_lnc_b", c_void_p),
                ("ca_wq", c_void_p), ("ca_wkv", c_void_p), ("ca_wo", c_void_p), ("ca_bo", c_void_p),
                ("ff_ln_w", c_void_p), ("ff_ln_b", c_void_p), ("ff_w1", c_void_p), ("ff_b1", c_void_p),
                ("ff_w2", c_void_p), ("ff_b2", c_void_p), ("w_stats", c_void_p), ("b_stats", c_void_p)]


class AeEncWorkspace(ctypes.Structure):
    _fields_ = [("max_points", i32), ("_pad", i32), ("feat", c_void_p), ("pe32", c_void_p), ("pe16", c_void_p),
                ("kbuf", c_void_p), ("vt", c_void_p), ("scores", c_void_p), ("prob", c_void_p), ("x", c_void_p),
                ("xq", c_void_p), ("att", c_void_p), ("xn", c_void_p), ("ff", c_void_p)]


class AeEncodeState:
    """Packed encoder-side weights + workspace of one KLAutoEncoder (kept on its AeRuntime)."""

    def __init__(self, rt):
        self.rt = rt
        self.sig = None
        self.ws_n = None

    def ensure_packed(self):
        rt = self.rt
        if self.sig == rt._sig:
            return
        m, dev, dim = rt.module, rt.device, rt.dim
        bf = torch.bfloat16
        keep = {}
        w = AeEncWeights()
        with torch.no_grad():
            def put(name, t, dtype):
                t = t.detach().to(dtype).contiguous()
                keep[name] = t
                setattr(w, name, t.data_ptr())
                return t
            w.dim, w.n_latents, w.latent_dim, w.heads = dim, m.num_latents, m.latent_dim, m.heads
            w.query_type = QUERY_TYPES[m.query_type]
            for i, v in enumerate(rt.freq24.tolist()):
                w.freq24[i] = v
            keep["wpe"] = rt.wpe_bf16
            w.wpe = rt.wpe_bf16.data_ptr()
            keep["pe_bias"] = rt.pe_bias
            w.pe_bias = rt.pe_bias.data_ptr()
            ca, cf = m.cross_attend_blocks
            put("ca_ln_w", ca.norm.weight, torch.float32); put("ca_ln_b", ca.norm.bias, torch.float32)
            put("ca_lnc_w", ca.norm_context.weight, torch.float32); put("ca_lnc_b", ca.norm_context.bias, torch.float32)
            put("ca_wq", ca.fn.to_q.weight, bf); put("ca_wkv", ca.fn.to_kv.weight, bf)
            put("ca_wo", ca.fn.to_out.weight, bf); put("ca_bo", ca.fn.to_out.bias, torch.float32)
            idx = geglu_pack_index(cf.fn.net[2].weight.shape[1], dev)
            put("ff_ln_w", cf.norm.weight, torch.float32); put("ff_ln_b", cf.norm.bias, torch.float32)
            put("ff_w1", cf.fn.net[0].weight[idx], bf); put("ff_b1", cf.fn.net[0].bias[idx], torch.float32)
            put("ff_w2", cf.fn.net[2].weight, bf); put("ff_b2", cf.fn.net[2].bias, torch.float32)
            if rt.deterministic:
                # AutoEncoder: no posterior head; the library copies the residual stream out (w_stats = NULL)
                w.stats_rows = dim
            else:
                L = m.latent_dim
                rows = -(-2 * L // 32) * 32
                ws_ = torch.zeros(rows, dim, device=dev, dtype=torch.float32)
                ws_[:L], ws_[L:2 * L] = m.mean_fc.weight.detach(), m.logvar_fc.weight.detach()
                bs_ = torch.zeros(rows, device=dev, dtype=torch.float32)
                bs_[:L], bs_[L:2 * L] = m.mean_fc.bias.detach(), m.logvar_fc.bias.detach()
                put("w_stats", ws_, bf); put("b_stats", bs_, torch.float32)
                w.stats_rows = rows
            if m.query_type == "learnable":
                put("latents", m.latents.weight, torch.float32)
            elif m.query_type == "mix":
                mx = m.mix_attn_layer
                put("mix_wkv", mx.fn.to_kv.weight, bf); put("mix_wo", mx.fn.to_out.weight, bf)
                put("mix_bo", mx.fn.to_out.bias, torch.float32)
                put("s_latents", m.s_latents.weight, torch.float32)
                put("wproj", m.query_proj.weight, bf); put("bproj", m.query_proj.bias, torch.float32)
                # to_q(LN(d_latents)) does not depend on the input: computed once here with the library's kernels
                st = _lib.cur_stream()
                dl = m.d_latents.weight.detach().float().contiguous()
                g, b = mx.norm.weight.detach().float().contiguous(), mx.norm.bias.detach().float().contiguous()
                dn = torch.empty(m.num_latents, dim, device=dev, dtype=bf)
                _lib.call("rald_ln_rows", dl.data_ptr(), dim, g.data_ptr(), b.data_ptr(), 0, 0, 0, dn.data_ptr(), dim, 0,
                          m.num_latents, dim, 1e-5, st)
                wq = mx.fn.to_q.weight.detach().to(bf).contiguous()
                q = torch.empty(m.num_latents, dim, device=dev, dtype=bf)
                _lib.call("rald_gemm_bf16", dn.data_ptr(), dim, wq.data_ptr(), dim, q.data_ptr(), dim, 0, 0, 0,
                          m.num_latents, dim, dim, 0, 0, st)
                torch.cuda.current_stream().synchronize()
                keep["mix_q"] = q
                w.mix_q = q.data_ptr()
        self.weights, self.keep, self.sig = w, keep, rt._sig

    def workspace(self, n: int) -> AeEncWorkspace:
        if self.ws_n != n:
            rt = self.rt
            dev, dim, M = rt.device, rt.dim, rt.module.num_latents
            npad = -(-n // 32) * 32
            bf = torch.bfloat16
            # zero-initialised: rows / columns [n, npad) are never written and must read as 0 (see ae_encode.cu)
            bufs = dict(feat=torch.zeros(npad, 64, device=dev, dtype=bf),
                        pe32=torch.zeros(npad, dim, device=dev, dtype=torch.float32),
                        pe16=torch.zeros(npad, dim, device=dev, dtype=bf),
                        kbuf=torch.zeros(npad, dim, device=dev, dtype=bf),
                        vt=torch.zeros(dim, npad, device=dev, dtype=bf),
                        scores=torch.zeros(M, npad, device=dev, dtype=torch.float32),
                        prob=torch.zeros(M, npad, device=dev, dtype=bf),
                        x=torch.zeros(M, dim, device=dev, dtype=torch.float32),
                        xq=torch.zeros(M, dim, device=dev, dtype=bf), att=torch.zeros(M, dim, device=dev, dtype=bf),
                        xn=torch.zeros(M, dim, device=dev, dtype=bf), ff=torch.zeros(M, 4 * dim, device=dev, dtype=bf))
            ws = AeEncWorkspace()
            ws.max_points = n
            for k, t in bufs.items():
                setattr(ws, k, t.data_ptr())
            self.ws, self.ws_bufs, self.ws_n = ws, bufs, n
        return self.ws


def _state(rt) -> AeEncodeState:
    st = rt.__dict__.get("_enc_state")
    if st is None:
        st = AeEncodeState(rt)
        rt.__dict__["_enc_state"] = st
    return st


def encode_raw(rt, pc: torch.Tensor):
    """pc [B, N, 3] -> (ml fp32 [B*M, stats_rows], fps_idx int64 [B, M] or None)."""
    if pc.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    es = _state(rt)
    es.ensure_packed()
    B, N, _ = pc.shape
    pc = pc.contiguous().float()
    M = rt.module.num_latents
    ml = torch.empty(B * M, es.weights.stats_rows, device=rt.device, dtype=torch.float32)
    idx = torch.empty(B, M, device=rt.device, dtype=torch.int64) if rt.module.query_type == "point" else None
    ws = es.workspace(N)
    _lib.call("rald_ae_encode_stats", ctypes.addressof(es.weights), ctypes.addressof(ws), pc.data_ptr(), B, N,
              _lib.ptr(idx), ml.data_ptr(), _lib.cur_stream())
    return ml, idx


def posterior(rt, ml: torch.Tensor, B: int, noise=None):
    """(mean, clamped logvar, z or None, kl [B]) from the packed (mean | logvar) rows."""
    M, L = rt.module.num_latents, rt.module.latent_dim
    dev = rt.device
    mean = torch.empty(B, M, L, device=dev, dtype=torch.float32)
    logvar = torch.empty_like(mean)
    z = torch.empty_like(mean) if noise is not None else None
    kl = torch.empty(B, device=dev, dtype=torch.float32)
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
    _lib.call("rald_ae_posterior", ml.data_ptr(), ml.shape[1], _lib.ptr(noise), B, M, L, mean.data_ptr(),
              logvar.data_ptr(), _lib.ptr(z), kl.data_ptr(), _lib.cur_stream())
    return mean, logvar, z, kl


def encode_stats(rt, pc: torch.Tensor):
    ml, _ = encode_raw(rt, pc)
    mean, logvar, _, _ = posterior(rt, ml, pc.shape[0])
    return mean, logvar
