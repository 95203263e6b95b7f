"""Drop-in for the reference's ``model/models_ae.py`` (VecSet KL autoencoder for LiDAR point clouds).

Same constructor signature, factories, parameter names / shapes / registration order and seeded-init RNG order
as the reference (331 state tensors for ``kl_d512_m512_l32_mix``). ``encode`` and
``decode`` run on librald_b200 (sm_100a): farthest point sampling, Fourier point embedding, long-KV
cross-attention, the 24-layer latent stack (the same tcgen05 GEMM / attention kernels as the denoiser) and a
streaming decoder-query kernel. ``decode`` caches the latent-stack output per latent tensor, because the
reference's ``evaluate`` decodes the same latents up to three times (engine_generation.py:204, 275, 300) and the
stack (116 GFLOP/frame) does not depend on the queries.

Reference: model/models_ae.py:34-49 (PreNorm), :51-68 (GEGLU FF), :70-105 (Attention), :108-138 (PointEmbed),
:141-179 (DiagonalGaussianDistribution), :284-432 (KLAutoEncoder), :434-512 (factories).
"""
from __future__ import annotations

from functools import wraps

import numpy as np
import torch
import torch.nn as nn


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def cache_fn(f):
    cache = None

    @wraps(f)
    def cached_fn(*args, _cache=True, **kwargs):
        if not _cache:
            return f(*args, **kwargs)
        nonlocal cache
        if cache is not None:
            return cache
        cache = f(*args, **kwargs)
        return cache
    return cached_fn


class _Identity(nn.Module):
    """Stand-in for timm's DropPath: the reference creates DropPath(0.1) modules (parameter-free, identity in
    eval). Stochastic depth only matters for training, which is outside the hot path."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


DropPath = _Identity


class PreNorm(nn.Module):
    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if exists(context_dim) else None


class GEGLU(nn.Module):
    pass


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, drop_path_rate=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), GEGLU(), nn.Linear(dim * mult, dim))
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()


class Attention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, drop_path_rate=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = default(context_dim, query_dim)
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, query_dim)
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()


class PointEmbed(nn.Module):
    def __init__(self, hidden_dim=48, dim=128):
        super().__init__()
        assert hidden_dim % 6 == 0
        self.embedding_dim = hidden_dim
        n = hidden_dim // 6
        e = torch.pow(2, torch.arange(n)).float() * np.pi
        z = torch.zeros(n)
        self.register_buffer("basis", torch.stack([torch.cat([e, z, z]), torch.cat([z, e, z]), torch.cat([z, z, e])]))
        self.mlp = nn.Linear(hidden_dim + 3, dim)


class DiagonalGaussianDistribution(object):
    """Posterior helper (reference :141-179); plain tensor arithmetic on whatever device mean/logvar live on."""

    def __init__(self, mean, logvar, deterministic=False):
        self.mean = mean
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        # the reference draws from the global CPU generator and moves the noise to the device (:153)
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.mean.device)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is None:
            return 0.5 * torch.mean(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2])
        return 0.5 * torch.mean(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                                - self.logvar + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.0])
        return 0.5 * torch.sum(np.log(2.0 * np.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)

    def mode(self):
        return self.mean


def _build_trunk(self, depth, dim, queries_dim, output_dim, heads, dim_head, weight_tie_layers, decoder_ff,
                 query_type, num_latents):
    """Registers the sub-modules shared by AutoEncoder / KLAutoEncoder in the reference's order."""
    self.cross_attend_blocks = nn.ModuleList([
        PreNorm(dim, Attention(dim, dim, heads=1, dim_head=dim), context_dim=dim),
        PreNorm(dim, FeedForward(dim)),
    ])
    self.point_embed = PointEmbed(dim=dim)
    latent_attn = cache_fn(lambda: PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, drop_path_rate=0.1)))
    latent_ff = cache_fn(lambda: PreNorm(dim, FeedForward(dim, drop_path_rate=0.1)))
    self.layers = nn.ModuleList([])
    tie = {"_cache": weight_tie_layers}
    self.query_type = query_type
    if query_type == "point":
        pass
    elif query_type == "learnable":
        self.latents = nn.Embedding(num_latents, dim)
    elif query_type == "mix":
        self.s_latents = nn.Embedding(num_latents, dim)
        self.d_latents = nn.Embedding(num_latents, dim)
        self.mix_attn_layer = latent_attn(**tie)
        self.query_proj = nn.Linear(dim, dim)
    else:
        raise NotImplementedError(f"Query type {query_type} is not implemented")
    for _ in range(depth):
        self.layers.append(nn.ModuleList([latent_attn(**tie), latent_ff(**tie)]))
    self.decoder_cross_attn = PreNorm(queries_dim, Attention(queries_dim, dim, heads=1, dim_head=dim), context_dim=dim)
    self.decoder_ff = PreNorm(queries_dim, FeedForward(queries_dim)) if decoder_ff else None
    self.to_outputs = nn.Linear(queries_dim, output_dim) if exists(output_dim) else nn.Identity()


class KLAutoEncoder(nn.Module):
    def __init__(self, *, depth=24, dim=512, queries_dim=512, output_dim=1, num_inputs=2048, num_latents=512,
                 latent_dim=64, heads=8, dim_head=64, weight_tie_layers=False, decoder_ff=False, query_type="point"):
        super().__init__()
        self.depth, self.num_inputs, self.num_latents = depth, num_inputs, num_latents
        self.dim, self.latent_dim, self.heads = dim, latent_dim, heads
        _build_trunk(self, depth, dim, queries_dim, output_dim, heads, dim_head, weight_tie_layers, decoder_ff,
                     query_type, num_latents)
        self.proj = nn.Linear(latent_dim, dim)
        self.mean_fc = nn.Linear(dim, latent_dim)
        self.logvar_fc = nn.Linear(dim, latent_dim)
        self.__dict__["_rt"] = None

    def _runtime(self):
        if self.__dict__.get("_rt") is None:
            from .runtime_ae import AeRuntime
            self.__dict__["_rt"] = AeRuntime(self)
        return self.__dict__["_rt"]

    @torch.no_grad()
    def encode_stats(self, pc):
        """(mean, clamped logvar) of the posterior, each [B, M, latent_dim] — encode() without the sampling."""
        B, N, D = pc.shape
        assert N == self.num_inputs
        return self._runtime().encode_stats(pc)

    def encode(self, pc):
        """pc [B, N, 3] -> (kl [B], z [B, M, latent_dim]) (reference :351-405)."""
        B, N, D = pc.shape
        assert N == self.num_inputs
        with torch.no_grad():
            # the reference draws the posterior noise from the global CPU generator and moves it over (:153)
            noise = torch.randn(B, self.num_latents, self.latent_dim)
            return self._runtime().encode(pc, noise)

    def decode(self, x, queries):
        """latents [B, M, latent_dim], queries [B, Q, 3] -> occupancy logits [B, Q, 1] (reference :408-424)."""
        with torch.no_grad():
            return self._runtime().decode(x, queries)

    def forward(self, pc, queries):
        kl, x = self.encode(pc)
        o = self.decode(x, queries).squeeze(-1)
        return {"logits": o, "kl": kl}


class AutoEncoder(nn.Module):
    """Deterministic variant ("not actually used" in the reference, :181-282): FPS queries, no posterior head, the
    latents are the 512-wide residual stream itself. Served by the same kernels as KLAutoEncoder (encode without the
    (mean | logvar) projection, decode without ``proj``); like there, dim 512 = 8 x 64 heads and 512 latents only
    (``ae_d512_m512``) — the other ``ae_d*_m*`` factories construct (state_dict contract) and raise on use."""

    def __init__(self, *, depth=24, dim=512, queries_dim=512, output_dim=1, num_inputs=2048, num_latents=512, heads=8,
                 dim_head=64, weight_tie_layers=False, decoder_ff=False):
        super().__init__()
        self.depth, self.num_inputs, self.num_latents = depth, num_inputs, num_latents
        self.dim, self.latent_dim, self.heads = dim, dim, heads
        _build_trunk(self, depth, dim, queries_dim, output_dim, heads, dim_head, weight_tie_layers, decoder_ff,
                     "point", num_latents)
        self.__dict__["_rt"] = None

    def _runtime(self):
        if self.__dict__.get("_rt") is None:
            from .runtime_ae import AeRuntime
            self.__dict__["_rt"] = AeRuntime(self)
        return self.__dict__["_rt"]

    def encode(self, pc):
        """pc [B, N, 3] -> latents [B, M, dim] (reference :226-257)."""
        B, N, D = pc.shape
        assert N == self.num_inputs
        with torch.no_grad():
            return self._runtime().encode_deterministic(pc)

    def decode(self, x, queries):
        """latents [B, M, dim], queries [B, Q, 3] -> occupancy logits [B, Q, 1] (reference :260-274)."""
        with torch.no_grad():
            return self._runtime().decode(x, queries)

    def forward(self, pc, queries):
        x = self.encode(pc)
        o = self.decode(x, queries).squeeze(-1)
        return {"logits": o}


def create_autoencoder(dim=512, M=512, latent_dim=64, N=2048, determinisitc=False, query_type="point"):
    if determinisitc:
        return AutoEncoder(depth=24, dim=dim, queries_dim=dim, output_dim=1, num_inputs=N, num_latents=M, heads=8,
                           dim_head=64)
    return KLAutoEncoder(depth=24, dim=dim, queries_dim=dim, output_dim=1, num_inputs=N, num_latents=M,
                         latent_dim=latent_dim, heads=8, dim_head=64, query_type=query_type)


def kl_d512_m512_l512(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=512, N=N)


def kl_d512_m512_l64(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=64, N=N)


def kl_d512_m512_l32(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=32, N=N)


def kl_d512_m512_l32_learn(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=32, N=N, query_type="learnable")


def kl_d512_m512_l32_mix(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=32, N=N, query_type="mix")


def kl_d512_m512_l16(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=16, N=N)


def kl_d512_m512_l8(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=8, N=N)


def kl_d512_m512_l4(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=4, N=N)


def kl_d512_m512_l2(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=2, N=N)


def kl_d512_m512_l1(N=2048):
    return create_autoencoder(dim=512, M=512, latent_dim=1, N=N)


def ae_d512_m512(N=2048):
    return create_autoencoder(dim=512, M=512, N=N, determinisitc=True)


def ae_d512_m256(N=2048):
    return create_autoencoder(dim=512, M=256, N=N, determinisitc=True)


def ae_d512_m128(N=2048):
    return create_autoencoder(dim=512, M=128, N=N, determinisitc=True)


def ae_d512_m64(N=2048):
    return create_autoencoder(dim=512, M=64, N=N, determinisitc=True)


def ae_d256_m512(N=2048):
    return create_autoencoder(dim=256, M=512, N=N, determinisitc=True)


def ae_d128_m512(N=2048):
    return create_autoencoder(dim=128, M=512, N=N, determinisitc=True)


def ae_d64_m512(N=2048):
    return create_autoencoder(dim=64, M=512, N=N, determinisitc=True)
