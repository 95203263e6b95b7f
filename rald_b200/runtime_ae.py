"""Host-side runtime of the VecSet autoencoder: weight packing / constant folding, workspaces, the decode-side
latent-stack cache, and the calls into the C ABI (include/rald_b200.h). No hot-path arithmetic happens here."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import c_void_p
from .runtime_dit import DitWorkspace, default_microbatch, geglu_pack_index, split_hi_lo


class AeWeights(ctypes.Structure):
    _fields_ = [("depth", ctypes.c_int32), ("dim", ctypes.c_int32), ("heads", ctypes.c_int32),
                ("latent_dim", ctypes.c_int32), ("n_latents", ctypes.c_int32), ("precise", ctypes.c_int32),
                ("w_qkv", c_void_p), ("w_o", c_void_p), ("w_ff1", c_void_p), ("w_ff2", c_void_p),
                ("b_o", c_void_p), ("b_ff1", c_void_p), ("b_ff2", c_void_p),
                ("ln1_w", c_void_p), ("ln1_b", c_void_p), ("ln2_w", c_void_p), ("ln2_b", c_void_p),
                ("proj_wt", c_void_p), ("proj_b", c_void_p)]


def precise_enabled() -> bool:
    """Split-weight (hi + lo bf16 pair) latent stack + K' fold with the erf GELU — the default. The 24-layer stack's
    weights rounded to plain bf16 shift the whole occupancy field by a common-mode offset (-2.5e-4 at
    random init, 17 % of the field's spatial standard deviation); RALD_B200_AE_PRECISE=0 restores plain bf16 weights
    and the logistic-form GELU (2.2 % less work per 64-frame step)."""
    return os.environ.get("RALD_B200_AE_PRECISE", "1") != "0"


class AeRuntime(_lib.RuntimeNotCopied):
    def __init__(self, module):
        self.module = module
        self._sig = None
        self._ws = None
        self._ws_frames = 0
        self._ctx_cache = None  # (z tensor, version, folded context)

    # ------------------------------------------------------------------ packing
    def _signature(self):
        ps = list(self.module.parameters())
        return (ps[0].device, sum(p._version for p in ps), len(ps), precise_enabled())

    def ensure_packed(self):
        sig = self._signature()
        if sig == self._sig:
            return
        m = self.module
        dev = sig[0]
        if dev.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only: move the module to a B200 (no CPU fallback)")
        dim, heads = m.dim, m.heads
        self.deterministic = not hasattr(m, "proj")     # AutoEncoder: latents of width dim, no proj / mean_fc / logvar_fc
        if dim != 512 or dim // heads != 64 or m.num_latents != 512:
            raise _lib.RaldError(f"unsupported autoencoder geometry dim={dim} heads={heads} latents={m.num_latents}: "
                                 "kernels are built for dim 512 = 8 x 64 and 512 latents")
        if m.decoder_ff is not None:
            raise _lib.RaldError("decoder_ff=True is not supported (the folded decoder needs the reference default)")
        bf = torch.bfloat16
        self.precise = precise_enabled()
        with torch.no_grad():
            layers = list(m.layers)
            def stack(fn, dtype):
                ws = [fn(a, f).detach() for a, f in layers]
                if dtype is bf and self.precise:   # weight matrices as split pairs [rows][2 cols]
                    return torch.stack([split_hi_lo(w) for w in ws]).contiguous()
                return torch.stack(ws).to(dtype).contiguous()
            idx = geglu_pack_index(layers[0][1].fn.net[2].weight.shape[1], dev)
            self.w_qkv = stack(lambda a, f: torch.cat([a.fn.to_q.weight, a.fn.to_kv.weight]), bf)
            self.w_o = stack(lambda a, f: a.fn.to_out.weight, bf)
            self.b_o = stack(lambda a, f: a.fn.to_out.bias, torch.float32)
            self.w_ff1 = stack(lambda a, f: f.fn.net[0].weight[idx], bf)
            self.b_ff1 = stack(lambda a, f: f.fn.net[0].bias[idx], torch.float32)
            self.w_ff2 = stack(lambda a, f: f.fn.net[2].weight, bf)
            self.b_ff2 = stack(lambda a, f: f.fn.net[2].bias, torch.float32)
            self.ln1_w = stack(lambda a, f: a.norm.weight, torch.float32)
            self.ln1_b = stack(lambda a, f: a.norm.bias, torch.float32)
            self.ln2_w = stack(lambda a, f: f.norm.weight, torch.float32)
            self.ln2_b = stack(lambda a, f: f.norm.bias, torch.float32)
            if self.deterministic:
                self.proj_wt = self.proj_b = None
            else:
                self.proj_wt = m.proj.weight.detach().float().t().contiguous()
                self.proj_b = m.proj.bias.detach().float().contiguous()
            # ---- decoder: constant folding in fp64 (weights only; see csrc/ae_query.cu) ----
            dca = m.decoder_cross_attn
            wq = dca.fn.to_q.weight.detach().double()
            wk, wv = dca.fn.to_kv.weight.detach().double().chunk(2, dim=0)
            wo, bo = dca.fn.to_out.weight.detach().double(), dca.fn.to_out.bias.detach().double()
            w_out, b_out = m.to_outputs.weight.detach().double(), m.to_outputs.bias.detach().double()
            if w_out.shape[0] != 1:
                raise _lib.RaldError("the folded decoder needs output_dim == 1 (every reference factory uses 1)")
            wf = (wq.t() @ wk).float()                                  # K' = LN_ctx(x) @ w_fold^T
            self.w_fold = split_hi_lo(wf) if self.precise else wf.to(bf).contiguous()
            u = (w_out @ wo)[0]                                          # [dim]
            self.w_vfold = (u @ wv).float().contiguous()                 # v' = LN_ctx(x) . w_vfold
            self.c0_scalar = float((w_out @ bo)[0] + b_out[0])
            self.ctx_ln_w = dca.norm_context.weight.detach().float().contiguous()
            self.ctx_ln_b = dca.norm_context.bias.detach().float().contiguous()
            self.q_ln_w = dca.norm.weight.detach().float().contiguous()
            self.q_ln_b = dca.norm.bias.detach().float().contiguous()
            pe = m.point_embed
            wpe = torch.zeros(dim, 64, device=dev, dtype=torch.float32)
            wpe[:, :pe.mlp.weight.shape[1]] = pe.mlp.weight.detach().float()
            self.wpe_bf16 = wpe.to(bf).contiguous()
            self.wpe_f32 = pe.mlp.weight.detach().float().contiguous()
            self.pe_bias = pe.mlp.bias.detach().float().contiguous()
            basis = pe.basis.detach().float().cpu()
            if basis.shape != (3, 24):
                raise _lib.RaldError(f"point_embed.basis has shape {tuple(basis.shape)}, expected (3, 24)")
            blocks = torch.stack([basis[a, 8 * a:8 * a + 8] for a in range(3)])
            off = basis.clone()
            for a in range(3):
                off[a, 8 * a:8 * a + 8] = 0
            if off.abs().max() != 0:
                raise _lib.RaldError("point_embed.basis is not block diagonal; the fused embedding assumes it is")
            self.freq24 = np.ascontiguousarray(blocks.numpy().reshape(-1), dtype=np.float32)
        self.device, self.dim = dev, dim
        self._ctx_cache = None
        self._sig = sig

    def _weights_struct(self) -> AeWeights:
        m = self.module
        w = AeWeights()
        w.depth, w.dim, w.heads = len(m.layers), m.dim, m.heads
        w.latent_dim, w.n_latents = m.latent_dim, m.num_latents
        w.precise = 1 if self.precise else 0
        for name in ("w_qkv", "w_o", "w_ff1", "w_ff2", "b_o", "b_ff1", "b_ff2", "ln1_w", "ln1_b", "ln2_w", "ln2_b",
                     "proj_wt", "proj_b"):
            setattr(w, name, _lib.ptr(getattr(self, name)) or None)
        return w

    def _workspace(self, frames: int):
        mb = max(1, min(default_microbatch(), frames))
        if self._ws is None or self._ws_frames != mb:
            T = mb * self.module.num_latents
            dev, dim = self.device, self.dim
            bufs = dict(h=torch.empty(1, device=dev), xn=torch.empty(T, dim, device=dev, dtype=torch.bfloat16),
                        qkv=torch.empty(T, 3 * dim, device=dev, dtype=torch.bfloat16),
                        att=torch.empty(T, dim, device=dev, dtype=torch.bfloat16),
                        ff=torch.empty(T, 4 * dim, device=dev, dtype=torch.bfloat16),
                        x_tmp=torch.empty(1, device=dev), d_tmp=torch.empty(1, device=dev))
            ws = DitWorkspace()
            ws.max_frames = mb
            for k, v in bufs.items():
                setattr(ws, k, v.data_ptr())
            self._ws, self._ws_bufs, self._ws_frames = ws, bufs, mb
        return self._ws

    # ------------------------------------------------------------------ decode
    def latent_stack(self, z: torch.Tensor) -> torch.Tensor:
        """z [B, M, latent_dim] -> stack output [B*M, dim] fp32."""
        self.ensure_packed()
        B, M, _ = z.shape
        z = z.contiguous().float()
        x = torch.empty(B * M, self.dim, device=self.device, dtype=torch.float32)
        w, ws = self._weights_struct(), self._workspace(B)
        _lib.call("rald_ae_stack", ctypes.addressof(w), ctypes.addressof(ws), z.data_ptr(), x.data_ptr(), B,
                  _lib.cur_stream())
        return x

    def _fold_context(self, x: torch.Tensor, B: int):
        """Per-frame decoder constants from the stack output: K' (bf16 [B*M, dim]), v' (fp32 [B, M]), c0 [B]."""
        rows = x.shape[0]
        st = _lib.cur_stream()
        cn = torch.empty(rows, self.dim, device=self.device, dtype=torch.bfloat16)
        _lib.call("rald_ln_rows", x.data_ptr(), self.dim, self.ctx_ln_w.data_ptr(), self.ctx_ln_b.data_ptr(), 0, 0, 0,
                  cn.data_ptr(), self.dim, 0, rows, self.dim, 1e-5, st)
        kp = torch.empty(rows, self.dim, device=self.device, dtype=torch.bfloat16)
        if self.precise:
            _lib.call("rald_gemm_bf16_wsplit", cn.data_ptr(), self.dim, self.w_fold.data_ptr(), 2 * self.dim,
                      kp.data_ptr(), self.dim, 0, 0, 0, rows, self.dim, self.dim, 0, 0, 0, 0, st)
        else:
            _lib.call("rald_gemm_bf16", cn.data_ptr(), self.dim, self.w_fold.data_ptr(), self.dim, kp.data_ptr(),
                      self.dim, 0, 0, 0, rows, self.dim, self.dim, 0, 0, st)
        vp = torch.empty(rows, device=self.device, dtype=torch.float32)
        _lib.call("rald_ln_dot_rows", x.data_ptr(), self.ctx_ln_w.data_ptr(), self.ctx_ln_b.data_ptr(),
                  self.w_vfold.data_ptr(), vp.data_ptr(), rows, self.dim, 1e-5, st)
        c0 = torch.full((B,), self.c0_scalar, device=self.device, dtype=torch.float32)
        return kp, vp, c0

    def context(self, z: torch.Tensor):
        """Folded decoder context of a latent set, cached per tensor object + version (the reference's evaluate()
        decodes the same latents up to three times: engine_generation.py:204, 275, 300)."""
        self.ensure_packed()
        if z.is_inference():
            # inference tensors carry no version counter: nothing to key the cache on, recompute
            return self._fold_context(self.latent_stack(z), z.shape[0])
        c = self._ctx_cache
        if c is not None and c[0] is z and c[1] == z._version:
            return c[2]
        x = self.latent_stack(z)
        ctx = self._fold_context(x, z.shape[0])
        self._ctx_cache = (z, z._version, ctx)
        return ctx

    def clear_cache(self):
        """Drops the cached latent stack / folded context. The cache is keyed on the latent tensor's identity and
        autograd version counter, which only torch ops bump: call this after refilling a latent tensor in place through
        a raw pointer (C ABI, DLPack consumer, ...)."""
        self._ctx_cache = None

    def query(self, ctx, queries: torch.Tensor) -> torch.Tensor:
        kp, vp, c0 = ctx
        B, Q, _ = queries.shape
        queries = queries.contiguous().float()
        logits = torch.empty(B, Q, device=self.device, dtype=torch.float32)
        if Q == 0:
            return logits
        _lib.call("rald_ae_query", queries.data_ptr(), B, Q, self.wpe_bf16.data_ptr(), self.pe_bias.data_ptr(),
                  self.q_ln_w.data_ptr(), self.q_ln_b.data_ptr(), kp.data_ptr(), vp.data_ptr(), c0.data_ptr(),
                  self.freq24.ctypes.data, logits.data_ptr(), self.dim, self.module.num_latents, _lib.cur_stream())
        return logits

    def decode(self, z: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
        if z.device.type != "cuda" or queries.device.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
        if z.shape[0] != queries.shape[0]:
            raise ValueError(f"batch mismatch: latents {z.shape[0]} vs queries {queries.shape[0]}")
        return self.query(self.context(z), queries).unsqueeze(-1)

    # ------------------------------------------------------------------ encode
    def _check_cuda(self):
        p = next(self.module.parameters())
        if p.device.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only: move the module to a B200 (no CPU fallback)")

    def encode_stats(self, pc: torch.Tensor):
        from .runtime_ae_encode import encode_stats
        self._check_cuda()
        self.ensure_packed()
        return encode_stats(self, pc)

    def encode_deterministic(self, pc: torch.Tensor) -> torch.Tensor:
        """AutoEncoder.encode (models_ae.py:226-257): pc [B, N, 3] -> latents [B, M, dim] fp32."""
        from .runtime_ae_encode import encode_raw
        self._check_cuda()
        self.ensure_packed()
        x, _ = encode_raw(self, pc)
        return x.view(pc.shape[0], self.module.num_latents, self.dim)

    def encode(self, pc: torch.Tensor, noise: torch.Tensor):
        """(kl [B], z [B, M, latent_dim]) with the posterior noise injected (drawn by the caller exactly as the
        reference does, from the global CPU generator: models_ae.py:153)."""
        from .runtime_ae_encode import encode_raw, posterior
        self._check_cuda()
        self.ensure_packed()
        ml, _ = encode_raw(self, pc)
        _, _, z, kl = posterior(self, ml, pc.shape[0], noise)
        return kl, z

    def fps(self, pc: torch.Tensor, m: int) -> torch.Tensor:
        """int64 [B, m] farthest-point indices into each cloud of pc [B, N, 3]."""
        if pc.device.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
        B, N, _ = pc.shape
        pc = pc.contiguous().float()
        idx = torch.empty(B, m, device=pc.device, dtype=torch.int64)
        _lib.call("rald_fps", pc.data_ptr(), B, N, m, idx.data_ptr(), _lib.cur_stream())
        return idx
