"""Training step of the denoiser on the sm_100a kernels: the differentiable ``LatentArrayTransformer`` behind
``EDMPrecond.forward`` when gradients are wanted (reference: ``EDMLoss.__call__``
model/models_radar_generation.py:277-295 under autograd, called from ``train_one_epoch`` engine_generation.py:89-110).

Shape of the computation (one ``torch.autograd.Function`` around the whole network):

* forward: the 24 blocks run as separate kernels — adaLN (``rald_ln_rows``), QKV projection, TMEM-resident attention
  that also emits its softmax statistics (``rald_attn_d64_stats``), output projection accumulated into the fp32
  residual stream by the GEMM's TMA reduce-add epilogue, the cross-attention against the 64 conditioning tokens, and
  the GEGLU feed-forward with its projection ``u`` materialised. Only the residual stream at every block input is
  kept (24 x [T, 512] fp32) — the reference checkpoints each ``BasicTransformerBlock`` the same way
  (``checkpoint=True``, model/models_radar_generation.py:143-146);
* backward: per block, in reverse, the block's forward is recomputed from its checkpoint and differentiated:
  dgrad GEMMs against transposed bf16 weight copies, wgrad GEMMs over K = rows with transposed activations
  (``rald_cast_transpose``) accumulating fp32 weight gradients, attention backward on tcgen05
  (``rald_attn_d64_bwd``), adaLN backward with per-frame scale / shift gradients (``rald_ln_bwd``), GEGLU backward,
  bias gradients by deterministic column sums. The timestep-embedding MLP and the 72 adaLN linears are differentiated
  from the per-frame modulation gradients.

Arithmetic: bf16 operands, fp32 accumulation, fp32 residual stream, gradient stream and weight gradients, erf GELU.
Everything here is host-side sequencing of C-ABI calls (include/rald_b200.h); there is no torch fallback for any
matrix product, attention or normalisation. Tiny [B, 512]-sized elementwise steps (SiLU and its derivative, the
cos / sin timestep features) are torch ops.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import _lib

BF = torch.bfloat16


def _p(t: Optional[torch.Tensor], byte_off: int = 0) -> int:
    return 0 if t is None else t.data_ptr() + byte_off


class DitTrainRuntime(_lib.RuntimeNotCopied):
    """bf16 weight copies (plain and transposed) of one EDMPrecond.model for the training step."""

    def __init__(self, module):
        self.module = module
        self._sig = None

    # ------------------------------------------------------------------ packing
    def _signature(self):
        ps = list(self.module.model.parameters())
        return (ps[0].device, sum(p._version for p in ps), len(ps))

    def ensure_packed(self):
        sig = self._signature()
        if sig == self._sig:
            return
        m = self.module.model
        dev = sig[0]
        if dev.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only: move the module to a B200 (no CPU fallback)")
        blocks = m.transformer_blocks
        dim = m.proj_in.weight.shape[0]
        heads = blocks[0].attn1.heads
        if dim != 512 or dim // heads != 64:
            raise _lib.RaldError(f"unsupported denoiser width {dim} / heads {heads}: kernels are built for 512 = 8 x 64")
        with torch.no_grad():
            def stack(fn, dtype=BF):
                return torch.stack([fn(b).detach() for b in blocks]).to(dtype).contiguous()

            def tr(w):   # [depth, N, K] -> [depth, K, N]: the W operand of the dgrad GEMM
                return w.transpose(1, 2).contiguous()
            self.w_qkv = stack(lambda b: torch.cat([b.attn1.to_q.weight, b.attn1.to_k.weight, b.attn1.to_v.weight]))
            self.w_o1 = stack(lambda b: b.attn1.to_out[0].weight)
            self.w_q2 = stack(lambda b: b.attn2.to_q.weight)
            self.w_kv2 = stack(lambda b: torch.cat([b.attn2.to_k.weight, b.attn2.to_v.weight]))
            self.w_o2 = stack(lambda b: b.attn2.to_out[0].weight)
            self.w_ff1 = stack(lambda b: b.ff.net[0].proj.weight)
            self.w_ff2 = stack(lambda b: b.ff.net[2].weight)
            for n in ("w_qkv", "w_o1", "w_q2", "w_kv2", "w_o2", "w_ff1", "w_ff2"):
                setattr(self, n + "_t", tr(getattr(self, n)))
            f32 = torch.float32
            self.b_o1 = stack(lambda b: b.attn1.to_out[0].bias, f32)
            self.b_o2 = stack(lambda b: b.attn2.to_out[0].bias, f32)
            self.b_ff1 = stack(lambda b: b.ff.net[0].proj.bias, f32)
            self.b_ff2 = stack(lambda b: b.ff.net[2].bias, f32)
            self.ada_w = torch.cat([torch.cat([b.norm1.linear.weight, b.norm2.linear.weight, b.norm3.linear.weight])
                                    .detach() for b in blocks]).float().contiguous()          # [depth*3*1024, 512]
            self.ada_b = torch.cat([torch.cat([b.norm1.linear.bias, b.norm2.linear.bias, b.norm3.linear.bias])
                                    .detach() for b in blocks]).float().contiguous()
            self.ada_w_t = self.ada_w.t().to(BF).contiguous()                                 # [512, depth*3*1024]
            self.map0_w = m.map_layer0.weight.detach().float().contiguous()
            self.map0_b = m.map_layer0.bias.detach().float().contiguous()
            self.map1_w = m.map_layer1.weight.detach().float().contiguous()
            self.map1_b = m.map_layer1.bias.detach().float().contiguous()
            self.ln_w = m.norm.weight.detach().float().contiguous()
            self.ln_b = m.norm.bias.detach().float().contiguous()
            C = m.proj_in.weight.shape[1]
            if C > 32:
                raise _lib.RaldError(f"latent channels {C} > 32 unsupported")
            self.proj_in_t = m.proj_in.weight.detach().float().t().contiguous()               # [C, 512]
            w_out = torch.zeros(32, dim, device=dev, dtype=torch.float32)
            w_out[:C] = m.proj_out.weight.detach().float()
            self.w_out = w_out.to(BF).contiguous()                                            # [32, 512]
            self.w_out_t = self.w_out.t().contiguous()                                        # [512, 32]
            half = m.t_channels // 2
            fr = torch.arange(half, dtype=torch.float32, device=dev) / half
            self.freqs = ((1.0 / m.map_noise.max_positions) ** fr).contiguous()
            self.half = half
            self.zero_bias = torch.zeros(dim, device=dev, dtype=torch.float32)
        self.depth, self.dim, self.heads, self.channels, self.device = len(blocks), dim, heads, C, dev
        self._sig = sig

    # ------------------------------------------------------------------ thin wrappers over the C ABI
    def _gemm(self, A, lda, W, ldw, out, ldo, M, N, K, out_mode=0, bias=None, resid=None, ldr=0):
        _lib.call("rald_gemm_bf16", A, lda, W, ldw, out, ldo, _p(bias), resid if isinstance(resid, int) else _p(resid),
                  ldr, M, N, K, out_mode, 0, _lib.cur_stream())

    def _gemm_f16cols(self, A, lda, W, ldw, out, ldo, M, N, K, f16_start, f16_period):
        _lib.call("rald_gemm_bf16_f16cols", A, lda, W, ldw, out, ldo, 0, M, N, K, f16_start, f16_period,
                  _lib.cur_stream())

    def _ln(self, x, gamma_ptr, beta_ptr, frame_stride, rows_per_frame, plus_one, out, rows):
        _lib.call("rald_ln_rows", x.data_ptr(), self.dim, gamma_ptr, beta_ptr, frame_stride, rows_per_frame, plus_one,
                  out.data_ptr(), self.dim, 0, rows, self.dim, 1e-5, _lib.cur_stream())

    def _transpose(self, t2d: torch.Tensor, want_plain: bool = False, colsum_out: Optional[torch.Tensor] = None):
        """[R, C] fp32 / bf16 (contiguous) -> (bf16 [R, C] or None, bf16 [C, R]); with colsum_out (fp32 [C]) the same pass
        also yields the column sums of the input (bias gradient)."""
        R, C = t2d.shape
        is_f32 = t2d.dtype == torch.float32
        plain = torch.empty(R, C, device=t2d.device, dtype=BF) if (want_plain and is_f32) else None
        Rp = (R + 7) // 8 * 8
        tt = torch.empty(C, Rp, device=t2d.device, dtype=BF) if Rp == R else torch.zeros(C, Rp, device=t2d.device, dtype=BF)
        fused = colsum_out is not None and R % 64 == 0 and C % 64 == 0
        part = torch.empty(R // 64, C, device=t2d.device, dtype=torch.float32) if fused else None
        _lib.call("rald_cast_transpose", t2d.data_ptr(), 1 if is_f32 else 0, C, R, C, _p(plain), C, tt.data_ptr(), Rp,
                  _p(part), _lib.cur_stream())
        if fused:
            _lib.call("rald_colsum_finish", part.data_ptr(), R // 64, C, colsum_out.data_ptr(), 0, _lib.cur_stream())
        elif colsum_out is not None:
            self._colsum(t2d, colsum_out)
        return (plain if is_f32 else (t2d if want_plain else None)), tt

    def _colsum(self, t2d: torch.Tensor, out: torch.Tensor):
        R, C = t2d.shape
        chunks = min(512, (R + 255) // 256)
        ws = torch.empty(chunks * C, device=t2d.device, dtype=torch.float32)
        _lib.call("rald_colsum", t2d.data_ptr(), 1 if t2d.dtype == torch.float32 else 0, C, R, C, ws.data_ptr(),
                  ws.numel(), out.data_ptr(), 0, _lib.cur_stream())

    def _wgrad(self, dyT: torch.Tensor, xT: torch.Tensor, out: torch.Tensor):
        """out fp32 [N_out, K_in] += dY^T X with dyT [N_out, T], xT [K_in, T] (bf16, K-major over the rows T)."""
        n_out, T = dyT.shape
        k_in = xT.shape[0]
        # `out` is zero-initialised: the K extent (rows of the batch) is split over several CTAs per output tile
        _lib.call("rald_gemm_bf16_accum", dyT.data_ptr(), T, xT.data_ptr(), xT.shape[1], out.data_ptr(), out.shape[-1],
                  n_out, k_in, T, _lib.cur_stream())

    def _sgemm(self, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta=0.0):
        _lib.call("rald_sgemm_f32", ta, tb, M, N, K, 1.0, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc,
                  _lib.cur_stream())

    # ------------------------------------------------------------------ timestep embedding
    def _temb_parts(self, sigma: torch.Tensor):
        """(e0 [B, 256], a0, s0, a1, t [B, 512]) of PositionalEmbedding + map_layer0/1 (:27-33, 217-219), fp32."""
        B = sigma.numel()
        c_noise = sigma.log() / 4
        arg = c_noise[:, None] * self.freqs[None, :]
        e0 = torch.cat([arg.cos(), arg.sin()], dim=1).contiguous()
        a0 = self.map0_b.repeat(B, 1)
        self._sgemm(0, 1, B, self.dim, 2 * self.half, e0, 2 * self.half, self.map0_w, 2 * self.half, a0, self.dim, beta=1.0)
        s0 = torch.nn.functional.silu(a0)
        a1 = self.map1_b.repeat(B, 1)
        self._sgemm(0, 1, B, self.dim, self.dim, s0, self.dim, self.map1_w, self.dim, a1, self.dim, beta=1.0)
        return e0, a0, s0, a1, torch.nn.functional.silu(a1)

    def _mod_table(self, sigma: torch.Tensor):
        B = sigma.numel()
        t_emb = torch.empty(B, self.dim, device=self.device, dtype=torch.float32)
        mod = torch.empty(B, self.depth, 3, 2 * self.dim, device=self.device, dtype=torch.float32)
        _lib.call("rald_dit_mod_table", sigma.data_ptr(), B, self.freqs.data_ptr(), self.half, self.map0_w.data_ptr(),
                  self.map0_b.data_ptr(), self.map1_w.data_ptr(), self.map1_b.data_ptr(), self.ada_w.data_ptr(),
                  self.ada_b.data_ptr(), self.depth, self.dim, t_emb.data_ptr(), mod.data_ptr(), _lib.cur_stream())
        return mod

    # ------------------------------------------------------------------ one block, forward (in place on h)
    def _block_forward(self, n: int, h: torch.Tensor, mod: torch.Tensor, tok16: torch.Tensor, B: int, M: int, L: int,
                       keep: Optional[Dict[str, torch.Tensor]] = None):
        dim, T, dev = self.dim, B * M, self.device
        fs = self.depth * 3 * 2 * dim                     # frame stride of the modulation table (floats)
        scale = 64 ** -0.5

        def mod_ptrs(i):
            base = mod.data_ptr() + ((n * 3 + i) * 2 * dim) * 4
            return base, base + dim * 4                   # scale, shift

        # ---- attn1 ----
        xn1 = torch.empty(T, dim, device=dev, dtype=BF)
        g, b_ = mod_ptrs(0)
        self._ln(h, g, b_, fs, M, 1, xn1, T)
        qkv = torch.empty(T, 3 * dim, device=dev, dtype=BF)      # V columns hold fp16 bit patterns
        self._gemm_f16cols(xn1.data_ptr(), dim, self.w_qkv[n].data_ptr(), dim, qkv.data_ptr(), 3 * dim, T, 3 * dim, dim,
                           2 * dim, 3 * dim)
        a1 = torch.empty(T, dim, device=dev, dtype=BF)
        st1 = torch.empty(T, self.heads, 2, device=dev, dtype=torch.float32)
        _lib.call("rald_attn_d64_stats", qkv.data_ptr(), 3 * dim, qkv.data_ptr() + dim * 2, 3 * dim,
                  qkv.data_ptr() + 2 * dim * 2, 3 * dim, a1.data_ptr(), dim, B, self.heads, M, M, scale, st1.data_ptr(),
                  _lib.cur_stream())
        self._gemm(a1.data_ptr(), dim, self.w_o1[n].data_ptr(), dim, h.data_ptr(), dim, T, dim, dim, out_mode=1,
                   bias=self.b_o1[n], resid=h, ldr=dim)
        h1 = h.clone() if keep is not None else None
        vb1 = vb2 = None
        if keep is not None:   # fp16 V columns, centred per frame and re-encoded as bf16 for the backward attention kernel
            vb1 = torch.empty(T, dim, device=dev, dtype=BF)
            _lib.call("rald_center_cast_f16_bf16", qkv.data_ptr() + 2 * dim * 2, 3 * dim, vb1.data_ptr(), dim, B, M, dim,
                      _lib.cur_stream())
        # ---- attn2 ----
        xn2 = torch.empty(T, dim, device=dev, dtype=BF)
        g, b_ = mod_ptrs(1)
        self._ln(h, g, b_, fs, M, 1, xn2, T)
        q2 = torch.empty(T, dim, device=dev, dtype=BF)
        self._gemm(xn2.data_ptr(), dim, self.w_q2[n].data_ptr(), dim, q2.data_ptr(), dim, T, dim, dim)
        kv2 = torch.empty(B * L, 2 * dim, device=dev, dtype=BF)  # K bf16 | V fp16
        self._gemm_f16cols(tok16.data_ptr(), dim, self.w_kv2[n].data_ptr(), dim, kv2.data_ptr(), 2 * dim, B * L, 2 * dim,
                           dim, dim, 2 * dim)
        a2 = torch.empty(T, dim, device=dev, dtype=BF)
        st2 = torch.empty(T, self.heads, 2, device=dev, dtype=torch.float32)
        _lib.call("rald_attn_d64_stats", q2.data_ptr(), dim, kv2.data_ptr(), 2 * dim, kv2.data_ptr() + dim * 2, 2 * dim,
                  a2.data_ptr(), dim, B, self.heads, M, L, scale, st2.data_ptr(), _lib.cur_stream())
        self._gemm(a2.data_ptr(), dim, self.w_o2[n].data_ptr(), dim, h.data_ptr(), dim, T, dim, dim, out_mode=1,
                   bias=self.b_o2[n], resid=h, ldr=dim)
        h2 = h.clone() if keep is not None else None
        if keep is not None:
            vb2 = torch.empty(B * L, dim, device=dev, dtype=BF)
            _lib.call("rald_center_cast_f16_bf16", kv2.data_ptr() + dim * 2, 2 * dim, vb2.data_ptr(), dim, B, L, dim,
                      _lib.cur_stream())
        # ---- feed-forward ----
        xn3 = torch.empty(T, dim, device=dev, dtype=BF)
        g, b_ = mod_ptrs(2)
        self._ln(h, g, b_, fs, M, 1, xn3, T)
        u = torch.empty(T, 8 * dim, device=dev, dtype=BF)
        self._gemm(xn3.data_ptr(), dim, self.w_ff1[n].data_ptr(), dim, u.data_ptr(), 8 * dim, T, 8 * dim, dim,
                   bias=self.b_ff1[n])
        gg = torch.empty(T, 4 * dim, device=dev, dtype=BF)
        _lib.call("rald_geglu_fwd", u.data_ptr(), T, 4 * dim, gg.data_ptr(), _lib.cur_stream())
        if keep is None:   # (the backward pass does not need h3)
            self._gemm(gg.data_ptr(), 4 * dim, self.w_ff2[n].data_ptr(), 4 * dim, h.data_ptr(), dim, T, dim, 4 * dim,
                       out_mode=1, bias=self.b_ff2[n], resid=h, ldr=dim)
        else:
            keep.update(xn1=xn1, qkv=qkv, a1=a1, st1=st1, h1=h1, vb1=vb1, vb2=vb2, xn2=xn2, q2=q2, kv2=kv2, a2=a2, st2=st2, h2=h2, xn3=xn3,
                        u=u, g=gg)

    # ------------------------------------------------------------------ forward
    def forward(self, x_in: torch.Tensor, sigma: torch.Tensor, tokens: torch.Tensor):
        """x_in fp32 [B, M, C] (already scaled by c_in), sigma fp32 [B], tokens fp32 [B, L, 512] ->
        (F fp32 [B, M, C], saved state for backward)."""
        self.ensure_packed()
        B, M, C = x_in.shape
        L = tokens.shape[1]
        if C != self.channels:
            raise ValueError(f"latents have {C} channels, the model {self.channels}")
        if M % 128 != 0:
            raise _lib.RaldError(f"training path needs n_latents % 128 == 0 (got {M})")
        if not (L == 64 or (L % 128 == 0 and L <= 512)):
            raise _lib.RaldError(f"training path supports 64 or 128 / 256 / 384 / 512 conditioning tokens (got {L})")
        T, dim, dev = B * M, self.dim, self.device
        x2d = x_in.reshape(T, C).contiguous().float()
        sigma = sigma.reshape(-1).contiguous().float()
        mod = self._mod_table(sigma)
        tok16 = tokens.reshape(B * L, dim).to(BF).contiguous()
        h = torch.empty(T, dim, device=dev, dtype=torch.float32)
        _lib.call("rald_linear_smallk", x2d.data_ptr(), C, self.proj_in_t.data_ptr(), self.zero_bias.data_ptr(),
                  h.data_ptr(), T, dim, _lib.cur_stream())
        ckpt = torch.empty(self.depth, T, dim, device=dev, dtype=torch.float32)
        for n in range(self.depth):
            ckpt[n].copy_(h)
            self._block_forward(n, h, mod, tok16, B, M, L)
        yn = torch.empty(T, dim, device=dev, dtype=BF)
        self._ln(h, self.ln_w.data_ptr(), self.ln_b.data_ptr(), 0, 0, 0, yn, T)
        F32 = torch.empty(T, 32, device=dev, dtype=torch.float32)
        self._gemm(yn.data_ptr(), dim, self.w_out.data_ptr(), dim, F32.data_ptr(), 32, T, 32, dim, out_mode=1)
        out = F32[:, :C].reshape(B, M, C).contiguous()
        saved = dict(B=B, M=M, L=L, C=C, x2d=x2d, sigma=sigma, mod=mod, tok16=tok16, ckpt=ckpt, h_final=h, yn=yn)
        return out, saved

    # ------------------------------------------------------------------ backward
    def backward(self, saved, dF: torch.Tensor):
        """dF fp32 [B, M, C] -> (dict parameter name (relative to EDMPrecond.model) -> fp32 gradient, dtokens fp32
        [B, L, 512])."""
        self.ensure_packed()
        B, M, L, C = saved["B"], saved["M"], saved["L"], saved["C"]
        T, dim, dev, depth, heads = B * M, self.dim, self.device, self.depth, self.heads
        f32 = torch.float32
        mod, tok16, ckpt = saved["mod"], saved["tok16"], saved["ckpt"]
        fs = depth * 3 * 2 * dim
        scale = 64 ** -0.5
        stream = _lib.cur_stream

        # stacked gradient buffers (views of them are returned per parameter)
        G = dict(w_qkv=torch.zeros(depth, 3 * dim, dim, device=dev, dtype=f32),
                 w_o1=torch.zeros(depth, dim, dim, device=dev, dtype=f32), b_o1=torch.empty(depth, dim, device=dev, dtype=f32),
                 w_q2=torch.zeros(depth, dim, dim, device=dev, dtype=f32),
                 w_kv2=torch.zeros(depth, 2 * dim, dim, device=dev, dtype=f32),
                 w_o2=torch.zeros(depth, dim, dim, device=dev, dtype=f32), b_o2=torch.empty(depth, dim, device=dev, dtype=f32),
                 w_ff1=torch.zeros(depth, 8 * dim, dim, device=dev, dtype=f32),
                 b_ff1=torch.empty(depth, 8 * dim, device=dev, dtype=f32),
                 w_ff2=torch.zeros(depth, dim, 4 * dim, device=dev, dtype=f32),
                 b_ff2=torch.empty(depth, dim, device=dev, dtype=f32))
        dmod = torch.empty(B, depth, 3, 2 * dim, device=dev, dtype=f32)
        dtok = torch.zeros(B * L, dim, device=dev, dtype=f32)
        ln_ws = torch.empty((T // 64) * 2 * dim, device=dev, dtype=f32)
        lse_ws = torch.empty(B * heads * M, device=dev, dtype=f32)
        ds_ws = torch.empty(B * heads * M, device=dev, dtype=f32)
        _, tokT = self._transpose(tok16)                                   # [512, B*L]

        def ln_bwd(x, dy, gamma_ptr, frame_stride, rows_per_frame, plus_one, dh, acc, dparam_ptr, gstride, wstride):
            _lib.call("rald_ln_bwd", x.data_ptr(), dy.data_ptr(), gamma_ptr, frame_stride, rows_per_frame, plus_one,
                      dh.data_ptr(), acc, ln_ws.data_ptr(), ln_ws.numel(), dparam_ptr, gstride, wstride, 0, T, dim, 1e-5,
                      stream())

        # ---- proj_out and the final LayerNorm ----
        dF32 = torch.zeros(T, 32, device=dev, dtype=f32)
        dF32[:, :C] = dF.reshape(T, C)
        dF16, dF16T = self._transpose(dF32, want_plain=True)               # [T, 32], [32, T]
        _, ynT = self._transpose(saved["yn"])                              # [512, T]
        g_wout = torch.zeros(32, dim, device=dev, dtype=f32)
        self._wgrad(dF16T, ynT, g_wout)
        dyn = torch.empty(T, dim, device=dev, dtype=BF)
        self._gemm(dF16.data_ptr(), 32, self.w_out_t.data_ptr(), 32, dyn.data_ptr(), dim, T, dim, 32)
        dh = torch.empty(T, dim, device=dev, dtype=f32)
        g_ln = torch.empty(2, dim, device=dev, dtype=f32)
        ln_bwd(saved["h_final"], dyn, self.ln_w.data_ptr(), 0, 0, 0, dh, 0, g_ln.data_ptr(), 0, dim)
        del dyn, ynT

        # ---- blocks, last to first ----
        h = torch.empty(T, dim, device=dev, dtype=f32)
        for n in reversed(range(depth)):
            k: Dict[str, torch.Tensor] = {}
            h.copy_(ckpt[n])
            self._block_forward(n, h, mod, tok16, B, M, L, keep=k)

            def mod_scale_ptr(i):
                return mod.data_ptr() + ((n * 3 + i) * 2 * dim) * 4

            def dmod_ptr(i):
                return dmod.data_ptr() + ((n * 3 + i) * 2 * dim) * 4

            # -- feed-forward: h3 = h2 + g W2^T + b2,  g = geglu(u),  u = xn3 W1^T + b1
            dh16, dh16T = self._transpose(dh, want_plain=True, colsum_out=G["b_ff2"][n])
            _, gT = self._transpose(k["g"])
            self._wgrad(dh16T, gT, G["w_ff2"][n])
            dg = torch.empty(T, 4 * dim, device=dev, dtype=BF)
            self._gemm(dh16.data_ptr(), dim, self.w_ff2_t[n].data_ptr(), dim, dg.data_ptr(), 4 * dim, T, 4 * dim, dim)
            du = torch.empty(T, 8 * dim, device=dev, dtype=BF)
            _lib.call("rald_geglu_bwd", k["u"].data_ptr(), dg.data_ptr(), T, 4 * dim, du.data_ptr(), stream())
            _, duT = self._transpose(du, colsum_out=G["b_ff1"][n])
            _, xn3T = self._transpose(k["xn3"])
            self._wgrad(duT, xn3T, G["w_ff1"][n])
            dxn = torch.empty(T, dim, device=dev, dtype=BF)
            self._gemm(du.data_ptr(), 8 * dim, self.w_ff1_t[n].data_ptr(), 8 * dim, dxn.data_ptr(), dim, T, dim, 8 * dim)
            ln_bwd(k["h2"], dxn, mod_scale_ptr(2), fs, M, 1, dh, 1, dmod_ptr(2), fs, dim)
            del gT, dg, du, duT, xn3T

            # -- attn2: h2 = h1 + a2 Wo2^T + bo2
            dh16, dh16T = self._transpose(dh, want_plain=True, colsum_out=G["b_o2"][n])
            _, a2T = self._transpose(k["a2"])
            self._wgrad(dh16T, a2T, G["w_o2"][n])
            da = torch.empty(T, dim, device=dev, dtype=BF)
            self._gemm(dh16.data_ptr(), dim, self.w_o2_t[n].data_ptr(), dim, da.data_ptr(), dim, T, dim, dim)
            dq2 = torch.empty(T, dim, device=dev, dtype=BF)
            dkv2 = torch.empty(B * L, 2 * dim, device=dev, dtype=BF)
            kv2 = k["kv2"]
            _lib.call("rald_attn_d64_bwd", k["q2"].data_ptr(), dim, kv2.data_ptr(), 2 * dim, k["vb2"].data_ptr(),
                      dim, da.data_ptr(), dim, k["st2"].data_ptr(), lse_ws.data_ptr(),
                      ds_ws.data_ptr(), dq2.data_ptr(), dim, dkv2.data_ptr(), 2 * dim, dkv2.data_ptr() + dim * 2, 2 * dim,
                      B, heads, M, L, scale, stream())
            _, dq2T = self._transpose(dq2)
            _, xn2T = self._transpose(k["xn2"])
            self._wgrad(dq2T, xn2T, G["w_q2"][n])
            self._gemm(dq2.data_ptr(), dim, self.w_q2_t[n].data_ptr(), dim, dxn.data_ptr(), dim, T, dim, dim)
            _, dkv2T = self._transpose(dkv2)                                 # [1024, B*L]
            self._wgrad(dkv2T, tokT, G["w_kv2"][n])
            self._gemm(dkv2.data_ptr(), 2 * dim, self.w_kv2_t[n].data_ptr(), 2 * dim, dtok.data_ptr(), dim, B * L, dim,
                       2 * dim, out_mode=1, resid=dtok, ldr=dim)
            ln_bwd(k["h1"], dxn, mod_scale_ptr(1), fs, M, 1, dh, 1, dmod_ptr(1), fs, dim)
            del a2T, dq2, dkv2, dq2T, xn2T, dkv2T

            # -- attn1: h1 = h0 + a1 Wo1^T + bo1
            dh16, dh16T = self._transpose(dh, want_plain=True, colsum_out=G["b_o1"][n])
            _, a1T = self._transpose(k["a1"])
            self._wgrad(dh16T, a1T, G["w_o1"][n])
            self._gemm(dh16.data_ptr(), dim, self.w_o1_t[n].data_ptr(), dim, da.data_ptr(), dim, T, dim, dim)
            dqkv = torch.empty(T, 3 * dim, device=dev, dtype=BF)
            qkv = k["qkv"]
            _lib.call("rald_attn_d64_bwd", qkv.data_ptr(), 3 * dim, qkv.data_ptr() + dim * 2, 3 * dim,
                      k["vb1"].data_ptr(), dim, da.data_ptr(), dim,
                      k["st1"].data_ptr(), lse_ws.data_ptr(), ds_ws.data_ptr(), dqkv.data_ptr(), 3 * dim,
                      dqkv.data_ptr() + dim * 2, 3 * dim, dqkv.data_ptr() + 2 * dim * 2, 3 * dim, B, heads, M, M, scale,
                      stream())
            _, dqkvT = self._transpose(dqkv)
            _, xn1T = self._transpose(k["xn1"])
            self._wgrad(dqkvT, xn1T, G["w_qkv"][n])
            self._gemm(dqkv.data_ptr(), 3 * dim, self.w_qkv_t[n].data_ptr(), 3 * dim, dxn.data_ptr(), dim, T, dim, 3 * dim)
            ln_bwd(ckpt[n], dxn, mod_scale_ptr(0), fs, M, 1, dh, 1, dmod_ptr(0), fs, dim)
            del k, a1T, dqkv, dqkvT, xn1T, da, dxn, dh16, dh16T

        # ---- proj_in: h0 = x_in Win^T ----
        _, dhT = self._transpose(dh)                                         # [512, T]
        x32 = torch.zeros(T, 32, device=dev, dtype=f32)
        x32[:, :C] = saved["x2d"]
        _, xT = self._transpose(x32)                                         # [32, T]
        g_win = torch.zeros(dim, 32, device=dev, dtype=f32)
        self._wgrad(dhT, xT, g_win)

        # ---- adaLN linears and the timestep-embedding MLP from dmod [B, depth*3*1024] ----
        R = depth * 3 * 2 * dim
        dmod2 = dmod.reshape(B, R)
        g_ada_b = torch.empty(R, device=dev, dtype=f32)
        self._colsum(dmod2, g_ada_b)
        e0, a0, s0, a1_, t = self._temb_parts(saved["sigma"])
        dmod16, dmodT = self._transpose(dmod2, want_plain=True)              # [B, R], [R, Bp]
        _, tT = self._transpose(t)                                           # [512, Bp]
        g_ada_w = torch.empty(R, dim, device=dev, dtype=f32)
        self._gemm(dmodT.data_ptr(), dmodT.shape[1], tT.data_ptr(), tT.shape[1], g_ada_w.data_ptr(), dim, R, dim,
                   dmodT.shape[1], out_mode=1)
        dt = torch.empty(B, dim, device=dev, dtype=f32)
        self._gemm(dmod16.data_ptr(), R, self.ada_w_t.data_ptr(), R, dt.data_ptr(), dim, B, dim, R, out_mode=1)

        def dsilu(x):
            s = torch.sigmoid(x)
            return s * (1 + x * (1 - s))
        da1 = (dt * dsilu(a1_)).contiguous()
        g_map1_w = torch.empty(dim, dim, device=dev, dtype=f32)
        self._sgemm(1, 0, dim, dim, B, da1, dim, s0, dim, g_map1_w, dim)
        ds0 = torch.empty(B, dim, device=dev, dtype=f32)
        self._sgemm(0, 0, B, dim, dim, da1, dim, self.map1_w, dim, ds0, dim)
        da0 = (ds0 * dsilu(a0)).contiguous()
        g_map0_w = torch.empty(dim, 2 * self.half, device=dev, dtype=f32)
        self._sgemm(1, 0, dim, 2 * self.half, B, da0, dim, e0, 2 * self.half, g_map0_w, 2 * self.half)
        g_map1_b = torch.empty(dim, device=dev, dtype=f32)
        g_map0_b = torch.empty(dim, device=dev, dtype=f32)
        self._colsum(da1, g_map1_b)
        self._colsum(da0, g_map0_b)

        # ---- per-parameter views ----
        grads: Dict[str, torch.Tensor] = {"proj_in.weight": g_win[:, :C], "norm.weight": g_ln[0], "norm.bias": g_ln[1],
                                          "proj_out.weight": g_wout[:C], "map_layer0.weight": g_map0_w,
                                          "map_layer0.bias": g_map0_b, "map_layer1.weight": g_map1_w,
                                          "map_layer1.bias": g_map1_b}
        ga = g_ada_w.reshape(depth, 3, 2 * dim, dim)
        gb = g_ada_b.reshape(depth, 3, 2 * dim)
        for n in range(depth):
            pre = f"transformer_blocks.{n}."
            grads[pre + "attn1.to_q.weight"] = G["w_qkv"][n, :dim]
            grads[pre + "attn1.to_k.weight"] = G["w_qkv"][n, dim:2 * dim]
            grads[pre + "attn1.to_v.weight"] = G["w_qkv"][n, 2 * dim:]
            grads[pre + "attn1.to_out.0.weight"] = G["w_o1"][n]
            grads[pre + "attn1.to_out.0.bias"] = G["b_o1"][n]
            grads[pre + "attn2.to_q.weight"] = G["w_q2"][n]
            grads[pre + "attn2.to_k.weight"] = G["w_kv2"][n, :dim]
            grads[pre + "attn2.to_v.weight"] = G["w_kv2"][n, dim:]
            grads[pre + "attn2.to_out.0.weight"] = G["w_o2"][n]
            grads[pre + "attn2.to_out.0.bias"] = G["b_o2"][n]
            grads[pre + "ff.net.0.proj.weight"] = G["w_ff1"][n]
            grads[pre + "ff.net.0.proj.bias"] = G["b_ff1"][n]
            grads[pre + "ff.net.2.weight"] = G["w_ff2"][n]
            grads[pre + "ff.net.2.bias"] = G["b_ff2"][n]
            for i in range(3):
                grads[pre + f"norm{i + 1}.linear.weight"] = ga[n, i]
                grads[pre + f"norm{i + 1}.linear.bias"] = gb[n, i]
        return grads, dtok.reshape(B, L, dim)


class DitTrainFunction(torch.autograd.Function):
    """F = LatentArrayTransformer(x_in, c_noise(sigma), tokens) with gradients for every parameter and the tokens."""

    @staticmethod
    def forward(ctx, rt: DitTrainRuntime, names: List[str], x_in, sigma, tokens, *params):
        with torch.no_grad():
            out, saved = rt.forward(x_in, sigma, tokens)
        ctx.rt, ctx.names, ctx.saved = rt, names, saved
        ctx.tokens_need_grad = tokens.requires_grad
        return out

    @staticmethod
    def backward(ctx, dF):
        with torch.no_grad():
            grads, dtok = ctx.rt.backward(ctx.saved, dF.contiguous().float())
        ctx.saved = None
        out = [grads[n] for n in ctx.names]
        return (None, None, None, None, dtok if ctx.tokens_need_grad else None, *out)


class RadarTokensFunction(torch.autograd.Function):
    """tokens = Linear(feat) + r / a / e position embeddings (process_radar_cond :390-405), differentiable in the
    projection, the embedding tables and the encoder features."""

    @staticmethod
    def forward(ctx, feat, w, b, r_emb, a_emb, e_emb):
        B, nr, na, ne, cz = feat.shape
        dim = w.shape[0]
        feat = feat.contiguous().float()
        tok = torch.empty(B, nr * na * ne, dim, device=feat.device, dtype=torch.float32)
        _lib.call("rald_radar_tokens", feat.data_ptr(), B, nr, na, ne, cz, w.data_ptr(), b.data_ptr(), r_emb.data_ptr(),
                  a_emb.data_ptr(), e_emb.data_ptr(), dim, tok.data_ptr(), 0, _lib.cur_stream())
        ctx.save_for_backward(feat, w)
        ctx.geom = (B, nr, na, ne, cz, dim, r_emb.shape[0], a_emb.shape[0], e_emb.shape[0])
        return tok

    @staticmethod
    def backward(ctx, dtok):
        feat, w = ctx.saved_tensors
        B, nr, na, ne, cz, dim, tr_, ta_, te_ = ctx.geom
        dev = feat.device
        dtok = dtok.contiguous().float()
        dw = torch.empty(dim, cz, device=dev, dtype=torch.float32)
        db = torch.empty(dim, device=dev, dtype=torch.float32)
        dr = torch.zeros(tr_, dim, device=dev, dtype=torch.float32)
        da = torch.zeros(ta_, dim, device=dev, dtype=torch.float32)
        de = torch.zeros(te_, dim, device=dev, dtype=torch.float32)
        _lib.call("rald_radar_tokens_bwd", dtok.data_ptr(), feat.data_ptr(), B, nr, na, ne, cz, dim, dw.data_ptr(),
                  db.data_ptr(), dr.data_ptr(), da.data_ptr(), de.data_ptr(), _lib.cur_stream())
        dfeat = None
        if ctx.needs_input_grad[0]:   # d feat = d tok W: into the (trainable) radar encoder
            wc = w.detach().float().contiguous()
            dfeat = torch.empty(B * nr * na * ne, cz, device=dev, dtype=torch.float32)
            _lib.call("rald_sgemm_f32", 0, 0, B * nr * na * ne, cz, dim, 1.0, dtok.data_ptr(), dim, wc.data_ptr(), cz, 0.0,
                      dfeat.data_ptr(), cz, _lib.cur_stream())
            dfeat = dfeat.reshape(B, nr, na, ne, cz)
        return dfeat, dw, db, dr, da, de
