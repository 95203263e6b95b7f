"""Drop-in for the reference's ``model/models_radar_generation.py`` (EDM-preconditioned latent-set denoiser).

Keeps the reference's public surface — ``EDMPrecond(n_latents, channels, ..., configs)``, ``forward(x, sigma,
label_tokens, cond_type)``, ``sample(cond, batch_seeds, cond_type)``, ``process_radar_cond``, ``edm_sampler``,
``StackedRandomGenerator``, ``EDMLoss`` and the eight ``kl_d512_m512_*_edm`` factories — and the exact
parameter names, shapes, registration order and seeded-init RNG order (637 tensors for the default config,
SURVEY.md Appendix A), so reference checkpoints load with ``strict=True``.

What differs is everything underneath: the network runs on the sm_100a kernels of librald_b200 through
``runtime_dit.DitRuntime`` (tcgen05 GEMMs with fused epilogues, TMEM-resident attention, fused
LayerNorm/adaLN passes, one fused precondition+Heun+projection kernel per evaluation), and work that the
reference repeats in every one of its 35 network evaluations is done once per ``sample()``: the radar encoder
and token embedding (reference :414-415), the cross-attention K/V projections of the tokens (:63-64) and the
timestep-embedding / adaLN linears (:217-219, :128-129).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .models_radar_encoder import Encoder as RadarEncoder
from .runtime_dit import DitRuntime
from .runtime_dit_train import DitTrainFunction, DitTrainRuntime, RadarTokensFunction
from .runtime_encoder_train import EncoderTrainFunction, EncoderTrainRuntime


def zero_module(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        p.detach().zero_()
    return module


# --------------------------------------------------------------------------------------------------
# parameter containers (names/order = checkpoint contract). Their forward passes are not used on the hot
# path: LatentArrayTransformer is evaluated as a whole by DitRuntime.
# --------------------------------------------------------------------------------------------------
class PositionalEmbedding(nn.Module):
    def __init__(self, num_channels, max_positions=10000, endpoint=False):
        super().__init__()
        self.num_channels, self.max_positions, self.endpoint = num_channels, max_positions, endpoint


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale, self.heads = dim_head ** -0.5, heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        if not glu:
            raise NotImplementedError("rald_b200: only the gated (GEGLU) feed-forward of the reference configs is built")
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out))


class AdaLayerNorm(nn.Module):
    def __init__(self, n_embd):
        super().__init__()
        self.silu = nn.SiLU()  # declared but never applied by the reference (:123-131); kept for parity
        self.linear = nn.Linear(n_embd, n_embd * 2)
        self.layernorm = nn.LayerNorm(n_embd, elementwise_affine=False)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1, self.norm2, self.norm3 = AdaLayerNorm(dim), AdaLayerNorm(dim), AdaLayerNorm(dim)
        self.checkpoint = checkpoint
        for i in (1, 2, 3):  # LayerScale(init 0) / DropPath(0) are identities in the reference (:146-163)
            setattr(self, f"ls{i}", nn.Identity())
            setattr(self, f"drop_path{i}", nn.Identity())


class LatentArrayTransformer(nn.Module):
    def __init__(self, in_channels, t_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None,
                 out_channels=None):
        super().__init__()
        self.in_channels, self.t_channels, self.context_dim = in_channels, t_channels, context_dim
        inner = n_heads * d_head
        self.proj_in = nn.Linear(in_channels, inner, bias=False)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim)
             for _ in range(depth)])
        self.norm = nn.LayerNorm(inner)
        if out_channels is None:
            self.proj_out = zero_module(nn.Linear(inner, in_channels, bias=False))
        else:
            self.num_cls = out_channels
            self.proj_out = zero_module(nn.Linear(inner, out_channels, bias=False))
        self.map_noise = PositionalEmbedding(t_channels)
        self.map_layer0 = nn.Linear(in_features=t_channels, out_features=inner)
        self.map_layer1 = nn.Linear(in_features=inner, out_features=inner)


# --------------------------------------------------------------------------------------------------
# sampler / RNG / loss
# --------------------------------------------------------------------------------------------------
class StackedRandomGenerator:
    """One torch.Generator per frame so a frame's noise depends only on its seed (reference :297-311)."""

    def __init__(self, device, seeds):
        super().__init__()
        self.generators = [torch.Generator(device).manual_seed(int(seed) % (1 << 32)) for seed in seeds]

    def randn(self, size, **kwargs):
        assert size[0] == len(self.generators)
        return torch.stack([torch.randn(size[1:], generator=gen, **kwargs) for gen in self.generators])

    def randn_like(self, input):
        return self.randn(input.shape, dtype=input.dtype, layout=input.layout, device=input.device)

    def randint(self, *args, size, **kwargs):
        assert size[0] == len(self.generators)
        return torch.stack([torch.randint(*args, size=size[1:], generator=gen, **kwargs) for gen in self.generators])


def karras_schedule(num_steps, sigma_min, sigma_max, rho, device="cpu") -> torch.Tensor:
    """t_i of reference :246-249 in fp32, with the trailing 0."""
    idx = torch.arange(num_steps, dtype=torch.float32, device=device)
    t = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def edm_sampler(net, latents, class_labels=None, cond_type=None, randn_like=torch.randn_like, num_steps=18,
                sigma_min=0.002, sigma_max=80, rho=7, S_churn=0, S_min=0, S_max=float("inf"), S_noise=1):
    """EDM 2nd-order Heun sampler (reference :235-275). With S_churn == 0 (the reference default) the whole loop
    — 2*num_steps-1 network evaluations and all scheduler arithmetic — is a single C call; with churn the
    per-step structure of the reference is kept and only the network evaluations are fused."""
    sigma_min = max(sigma_min, net.sigma_min)
    sigma_max = min(sigma_max, net.sigma_max)
    t_steps = karras_schedule(num_steps, sigma_min, sigma_max, rho)
    t_steps = torch.cat([net.round_sigma(t_steps[:-1]), t_steps[-1:]])
    tokens_bf16 = net._condition(class_labels, cond_type)
    if S_churn == 0:
        return net._runtime().sample(latents, tokens_bf16, t_steps)
    # stochastic variant: noise injection between steps needs the host-visible loop
    x_next = latents.to(torch.float32) * t_steps[0]
    for i in range(num_steps):
        t_cur, t_next = t_steps[i], t_steps[i + 1]
        gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= t_cur <= S_max else 0
        t_hat = net.round_sigma(t_cur + gamma * t_cur)
        x_hat = x_next + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * randn_like(x_next)
        denoised = net._runtime().forward(x_hat, t_hat, tokens_bf16)
        d_cur = (x_hat - denoised) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            denoised = net._runtime().forward(x_next, t_next, tokens_bf16)
            d_prime = (x_next - denoised) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next


class EDMLoss:
    """EDM denoising loss of the reference (:277-295): per-sample log-normal sigma, weight (s^2 + sd^2) / (s sd)^2,
    mean of weight * (D(y + n; sigma) - y)^2. The RNG calls (``randn([B,1,1])``, then ``randn_like(y)``, both on the
    input's device) are the reference's, in its order. With a ``.train()`` network whose parameters require gradients
    (and gradients enabled) the network evaluation is the differentiable training forward of
    ``runtime_dit_train`` and ``loss.backward()`` fills the ``.grad`` of every denoiser parameter
    (engine_generation.py:89-110); otherwise the inference kernels evaluate it (validation loss)."""

    def __init__(self, P_mean=-1.2, P_std=1.2, sigma_data=1):
        self.P_mean, self.P_std, self.sigma_data = P_mean, P_std, sigma_data

    def __call__(self, net, inputs, labels=None, cond_type=None, augment_pipe=None):
        rnd_normal = torch.randn([inputs.shape[0], 1, 1], device=inputs.device)
        sigma = (rnd_normal * self.P_std + self.P_mean).exp()
        weight = (sigma ** 2 + self.sigma_data ** 2) / (sigma * self.sigma_data) ** 2
        y, _ = augment_pipe(inputs) if augment_pipe is not None else (inputs, None)
        n = torch.randn_like(y) * sigma
        D_yn = net(y + n, sigma, labels, cond_type)
        loss = weight * ((D_yn - y) ** 2)
        return loss.mean()


# --------------------------------------------------------------------------------------------------
# EDMPrecond
# --------------------------------------------------------------------------------------------------
class EDMPrecond(nn.Module):
    def __init__(self, n_latents=512, channels=8, use_fp16=False, sigma_min=0, sigma_max=float("inf"), sigma_data=1,
                 n_heads=8, d_head=64, depth=12, configs=None):
        super().__init__()
        self.n_latents, self.channels, self.use_fp16 = n_latents, channels, use_fp16
        self.sigma_min, self.sigma_max, self.sigma_data = sigma_min, sigma_max, sigma_data
        self.configs = configs
        self.model = LatentArrayTransformer(in_channels=channels, t_channels=256, n_heads=n_heads, d_head=d_head,
                                            depth=depth)
        self.unfreeze_radar_enc = self.configs.get("unfreeze_radar_enc", False)
        if configs.cond_type == "radar":
            self.radar_token_channel = self.configs.radar_token_channel
            if self.unfreeze_radar_enc:
                self.radar_enc = RadarEncoder(in_channels=1, ch=self.configs.enc_hidden_ch,
                                              z_channels=self.configs.enc_radar_ch)
            pre = "enc_radar" if self.configs.use_radar_enc else "input_radar"
            tc = self.radar_token_channel
            self.radar_r_emb = nn.Embedding(self.configs[pre + "_r_dim"], tc)
            self.radar_a_emb = nn.Embedding(self.configs[pre + "_a_dim"], tc)
            self.radar_e_emb = nn.Embedding(self.configs[pre + "_e_dim"], tc)
            self.radar_token_project = nn.Linear(self.configs.enc_radar_ch if self.configs.use_radar_enc else 1, tc)
        self.__dict__["_rt"] = None
        self.__dict__["_trt"] = None
        self.__dict__["_ert"] = None

    # ---- runtime plumbing -------------------------------------------------------------------------
    def _runtime(self) -> DitRuntime:
        if self.__dict__.get("_rt") is None:
            self.__dict__["_rt"] = DitRuntime(self)
        return self.__dict__["_rt"]

    def _tokens(self, radar_cube: torch.Tensor, want_f32: bool, want_bf16: bool):
        """Radar cube [B, R, A, E, ch] -> conditioning tokens (reference process_radar_cond :363-407)."""
        if radar_cube.device.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
        feat = radar_cube[..., 0:1].contiguous().float()  # intensity only (:378)
        if self.configs.get("unfreeze_radar_enc", False):
            feat = self.radar_enc.forward_channels_last(feat)  # [B, r, a, e, cz]
        B, nr, na, ne, cz = feat.shape
        for emb, n, nm in ((self.radar_r_emb, nr, "range"), (self.radar_a_emb, na, "azimuth"),
                           (self.radar_e_emb, ne, "elevation")):
            if n > emb.weight.shape[0]:
                raise IndexError(f"{nm} dimension {n} exceeds the embedding table size {emb.weight.shape[0]}")
        dim = self.radar_token_channel
        dev = feat.device
        p = self.radar_token_project
        if cz != p.in_features:
            # the reference fails here with a matmul shape error (F.linear of [.., cz] with a [dim, in_features] weight)
            raise ValueError(f"radar_token_project expects {p.in_features} input channels, the conditioning has {cz}")
        for nm, t in (("radar_token_project.weight", p.weight), ("radar_token_project.bias", p.bias),
                      ("radar_r_emb.weight", self.radar_r_emb.weight), ("radar_a_emb.weight", self.radar_a_emb.weight),
                      ("radar_e_emb.weight", self.radar_e_emb.weight)):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise _lib.RaldError(f"{nm} must be a contiguous fp32 tensor on {dev} (got {t.dtype}, "
                                     f"contiguous={t.is_contiguous()}, {t.device})")
        tok32 = torch.empty(B, nr * na * ne, dim, device=dev, dtype=torch.float32) if want_f32 else None
        tok16 = torch.empty(B * nr * na * ne, dim, device=dev, dtype=torch.bfloat16) if want_bf16 else None
        _lib.call("rald_radar_tokens", feat.data_ptr(), B, nr, na, ne, cz, p.weight.data_ptr(), p.bias.data_ptr(),
                  self.radar_r_emb.weight.data_ptr(), self.radar_a_emb.weight.data_ptr(),
                  self.radar_e_emb.weight.data_ptr(), dim, _lib.ptr(tok32), _lib.ptr(tok16), _lib.cur_stream())
        return tok32, tok16

    def _condition(self, label_tokens, cond_type) -> torch.Tensor:
        """bf16 conditioning tokens [B*L, dim]; accepts a radar cube or (private) precomputed tokens [B, L, dim]."""
        if cond_type != "radar":
            raise ValueError(f"cond_type={cond_type!r}: only 'radar' conditioning exists in the reference forward")
        if label_tokens.dim() == 3:  # already tokens
            return label_tokens.reshape(-1, label_tokens.shape[-1]).to(torch.bfloat16).contiguous()
        return self._tokens(label_tokens, want_f32=False, want_bf16=True)[1]

    # ---- reference API ----------------------------------------------------------------------------
    @torch.no_grad()
    def process_radar_cond(self, radar_cube):
        return self._tokens(radar_cube, want_f32=True, want_bf16=False)[0]

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())

    def _train_runtime(self) -> DitTrainRuntime:
        if self.__dict__.get("_trt") is None:
            self.__dict__["_trt"] = DitTrainRuntime(self)
        return self.__dict__["_trt"]

    def _enc_train_runtime(self) -> EncoderTrainRuntime:
        if self.__dict__.get("_ert") is None:
            self.__dict__["_ert"] = EncoderTrainRuntime(self.radar_enc)
        return self.__dict__["_ert"]

    def _tokens_train(self, label_tokens, cond_type) -> torch.Tensor:
        """fp32 conditioning tokens [B, L, dim] for the training forward, differentiable in radar_token_project, the
        r / a / e embeddings and — when its parameters require gradients — the radar encoder
        (runtime_encoder_train.EncoderTrainFunction); a frozen encoder runs on the inference kernels (the reference's
        `radar_enc._encode` under no_grad, engine_generation.py:86-87)."""
        if cond_type != "radar":
            raise ValueError(f"cond_type={cond_type!r}: only 'radar' conditioning exists in the reference forward")
        if label_tokens.dim() == 3:  # already tokens
            return label_tokens.to(torch.float32)
        if label_tokens.device.type != "cuda":
            raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
        feat = label_tokens[..., 0:1].contiguous().float()
        if self.configs.get("unfreeze_radar_enc", False):
            if any(p.requires_grad for p in self.radar_enc.parameters()):
                # trained jointly with the denoiser (the shipped configuration): differentiable encoder
                named = list(self.radar_enc.named_parameters())
                feat = EncoderTrainFunction.apply(self._enc_train_runtime(), [n for n, _ in named], feat,
                                                  *[p for _, p in named])
            else:
                with torch.no_grad():
                    feat = self.radar_enc.forward_channels_last(feat)
        p = self.radar_token_project
        if feat.shape[-1] != p.in_features:
            raise ValueError(f"radar_token_project expects {p.in_features} input channels, the conditioning has "
                             f"{feat.shape[-1]}")
        return RadarTokensFunction.apply(feat, p.weight, p.bias, self.radar_r_emb.weight, self.radar_a_emb.weight,
                                         self.radar_e_emb.weight)

    def _forward_train(self, x, sigma, label_tokens, cond_type):
        """The reference's forward (:412-430) with the network evaluation as ONE autograd node on the sm_100a
        kernels (runtime_dit_train.DitTrainFunction); the preconditioning arithmetic around it is the reference's."""
        tokens = self._tokens_train(label_tokens, cond_type)
        x = x.to(torch.float32)
        sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device).reshape(-1, 1, 1)
        if sigma.shape[0] == 1 and x.shape[0] > 1:
            sigma = sigma.expand(x.shape[0], 1, 1)
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + sigma ** 2).sqrt()
        named = [(n, p) for n, p in self.model.named_parameters()]
        F_x = DitTrainFunction.apply(self._train_runtime(), [n for n, _ in named], (c_in * x).contiguous(),
                                     sigma.flatten().contiguous(), tokens, *[p for _, p in named])
        return c_skip * x + c_out * F_x

    def forward(self, x, sigma, label_tokens=None, cond_type=None, force_fp32=False, **model_kwargs):
        """D_x = c_skip x + c_out F(c_in x, ln(sigma)/4, cond) (reference :412-430), fp32 in / fp32 out.
        sigma: 0-d, [B] or [B,1,1]. In .train() mode with gradients enabled and trainable parameters this is the
        differentiable training forward; otherwise the inference kernels run and nothing is recorded."""
        if self._wants_grad():
            return self._forward_train(x, sigma, label_tokens, cond_type)
        with torch.no_grad():
            tokens = self._condition(label_tokens, cond_type)
            sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device)
            return self._runtime().forward(x.to(torch.float32), sigma, tokens)

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)

    @torch.no_grad()
    def sample(self, cond, batch_seeds=None, cond_type=None):
        if cond is not None:
            batch_size, device = cond.shape[0], cond.device
            if batch_seeds is None:
                batch_seeds = torch.arange(batch_size)
        else:
            device = batch_seeds.device
            batch_size = batch_seeds.shape[0]
        rnd = StackedRandomGenerator(device, batch_seeds)
        latents = rnd.randn([batch_size, self.n_latents, self.channels], device=device)
        return edm_sampler(self, latents, cond, cond_type, randn_like=rnd.randn_like)

    @torch.no_grad()
    def sample_from_latents(self, latents, cond, cond_type="radar", num_steps=18, trace=False):
        """Same as sample() with the initial unit-normal latents injected (parity tests draw them on the CPU,
        where the reference oracle runs). Returns x, or (x, per-step x_next [num_steps, B, M, C]) with trace."""
        t_steps = karras_schedule(num_steps, max(0.002, self.sigma_min), min(80, self.sigma_max), 7)
        tokens = self._condition(cond, cond_type)
        tr = torch.empty(num_steps, *latents.shape, device=latents.device, dtype=torch.float32) if trace else None
        out = self._runtime().sample(latents, tokens, t_steps, trace=tr)
        return (out, tr) if trace else out


def kl_d512_m512_l8_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=8, configs=configs)


def kl_d512_m512_l16_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=16, configs=configs)


def kl_d512_m512_l32_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=32, configs=configs)


def kl_d512_m512_l4_d24_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=4, depth=24, configs=configs)


def kl_d512_m512_l8_d24_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=8, depth=24, configs=configs)


def kl_d512_m512_l32_d24_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=32, depth=24, configs=configs)


def kl_d512_m512_l32_d18_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=32, depth=18, configs=configs)


def kl_d512_m512_l32_d12_edm(configs=None):
    return EDMPrecond(n_latents=512, channels=32, depth=12, configs=configs)
