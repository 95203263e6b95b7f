"""CPU fp32 restatement of the RaLD generation hot path — TEST INFRASTRUCTURE ONLY.

This module is the parity oracle for rald_b200. It restates, as plain functions over a reference-layout
``state_dict`` (fp32 torch tensors on the CPU), the algorithms of

    model/models_radar_generation.py   (EDM preconditioning, DiT-style latent-set denoiser, Heun sampler)
    model/models_radar_encoder.py      (3-D conv radar-cube encoder)
    model/models_ae.py                 (VecSet KL autoencoder encode / decode, point embedding, FPS call)

of RoyAPTX4869/RaLD (paths relative to the reference root; every function cites the lines it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import it; the
product (``rald_b200``) never does. Parity status: PINNED for everything except FPS — the functions are
checked against the unmodified reference modules (tests/golden/make_golden.py imports them from
/root/reference and commits the outputs under tests/golden/). FPS restates torch_cluster==1.6.3's published
algorithm with start index 0 (the library itself is not in the reference tree): "parity unpinned" for FPS.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# shared building blocks
# ----------------------------------------------------------------------------------------------
def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _heads_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v per head; q [B,Sq,h*d], k/v [B,Skv,h*d].
    models_radar_generation.py:66-75 and models_ae.py:91-104 (identical arithmetic)."""
    B, Sq, D = q.shape
    d = D // heads
    qh = q.view(B, Sq, heads, d).transpose(1, 2)
    kh = k.view(B, -1, heads, d).transpose(1, 2)
    vh = v.view(B, -1, heads, d).transpose(1, 2)
    sim = torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.matmul(attn, vh)
    return out.transpose(1, 2).reshape(B, Sq, D)


def _geglu_ff(x: torch.Tensor, w1, b1, w2, b2) -> torch.Tensor:
    """Linear -> (value, gate) halves -> value * gelu_erf(gate) -> Linear.
    models_radar_generation.py:88-117, models_ae.py:51-68."""
    h = F.linear(x, w1, b1)
    val, gate = h.chunk(2, dim=-1)
    return F.linear(val * F.gelu(gate), w2, b2)


# ----------------------------------------------------------------------------------------------
# denoiser (models_radar_generation.py)
# ----------------------------------------------------------------------------------------------
def positional_embedding(x: torch.Tensor, num_channels: int = 256, max_positions: int = 10000) -> torch.Tensor:
    """models_radar_generation.py:27-33 — cos first, then sin; freqs = max_positions^(-j/half)."""
    half = num_channels // 2
    freqs = torch.arange(half, dtype=torch.float32, device=x.device) / half
    freqs = (1.0 / max_positions) ** freqs
    y = torch.outer(x.to(torch.float32), freqs)
    return torch.cat([y.cos(), y.sin()], dim=1)


def timestep_embedding(sd: SD, c_noise: torch.Tensor, prefix: str = "model.") -> torch.Tensor:
    """models_radar_generation.py:217-219 -> [S,1,512]."""
    t = positional_embedding(c_noise)[:, None]
    t = F.silu(_lin(sd, prefix + "map_layer0", t))
    return F.silu(_lin(sd, prefix + "map_layer1", t))


def ada_layer_norm(sd: SD, name: str, x: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
    """models_radar_generation.py:127-131 — no SiLU; scale is the first half of the linear output."""
    emb = _lin(sd, name + ".linear", t_emb)
    scale, shift = emb.chunk(2, dim=2)
    return F.layer_norm(x, (x.shape[-1],)) * (1 + scale) + shift


def _dit_attention(sd: SD, name: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """models_radar_generation.py:55-76."""
    ctx = x if context is None else context
    q = F.linear(x, sd[name + ".to_q.weight"])
    k = F.linear(ctx, sd[name + ".to_k.weight"])
    v = F.linear(ctx, sd[name + ".to_v.weight"])
    o = _heads_attention(q, k, v, heads)
    return F.linear(o, sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])


def dit_depth(sd: SD, prefix: str = "model.") -> int:
    n = 0
    while f"{prefix}transformer_blocks.{n}.attn1.to_q.weight" in sd:
        n += 1
    return n


def dit_forward(sd: SD, x: torch.Tensor, c_noise: torch.Tensor, cond: torch.Tensor, heads: int = 8,
                prefix: str = "model.") -> torch.Tensor:
    """LatentArrayTransformer.forward, models_radar_generation.py:215-233; block at :165-169."""
    t_emb = timestep_embedding(sd, c_noise, prefix)
    h = F.linear(x, sd[prefix + "proj_in.weight"])
    for n in range(dit_depth(sd, prefix)):
        b = f"{prefix}transformer_blocks.{n}."
        h = _dit_attention(sd, b + "attn1", ada_layer_norm(sd, b + "norm1", h, t_emb), None, heads) + h
        h = _dit_attention(sd, b + "attn2", ada_layer_norm(sd, b + "norm2", h, t_emb), cond, heads) + h
        h = _geglu_ff(ada_layer_norm(sd, b + "norm3", h, t_emb), sd[b + "ff.net.0.proj.weight"],
                      sd[b + "ff.net.0.proj.bias"], sd[b + "ff.net.2.weight"], sd[b + "ff.net.2.bias"]) + h
    h = F.layer_norm(h, (h.shape[-1],), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return F.linear(h, sd[prefix + "proj_out.weight"])


def edm_precond(sd: SD, x: torch.Tensor, sigma: torch.Tensor, cond_tokens: torch.Tensor, sigma_data: float = 1.0,
                heads: int = 8) -> torch.Tensor:
    """EDMPrecond.forward after conditioning, models_radar_generation.py:418-430."""
    x = x.to(torch.float32)
    sigma = torch.as_tensor(sigma, dtype=torch.float32).to(x.device).reshape(-1, 1, 1)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.log() / 4
    f_x = dit_forward(sd, c_in * x, c_noise.flatten(), cond_tokens, heads)
    return c_skip * x + c_out * f_x


def karras_sigmas(num_steps: int = 18, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0
                  ) -> torch.Tensor:
    """models_radar_generation.py:246-249 — fp32 arithmetic, trailing 0."""
    idx = torch.arange(num_steps, dtype=torch.float32)
    t = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def stacked_randn(seeds: Sequence[int], shape: Sequence[int]) -> torch.Tensor:
    """StackedRandomGenerator('cpu', seeds).randn([B,*shape]), models_radar_generation.py:297-304."""
    outs = []
    for s in seeds:
        g = torch.Generator("cpu").manual_seed(int(s) % (1 << 32))
        outs.append(torch.randn(list(shape), generator=g))
    return torch.stack(outs)


def edm_sample(sd: SD, latents: torch.Tensor, cond_tokens: torch.Tensor, num_steps: int = 18,
               sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0, heads: int = 8,
               trace: Optional[list] = None, stop_after: Optional[int] = None, S_churn: float = 0.0,
               S_min: float = 0.0, S_max: float = float("inf"), S_noise: float = 1.0,
               noises: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """edm_sampler, models_radar_generation.py:235-275. With S_churn = 0 (the reference default) gamma = 0 and
    x_hat = x_cur; with churn (:254-260) the per-step noise randn_like(x_cur) is injected through `noises` (one
    tensor per step, drawn by the caller in the reference's order).
    `cond_tokens` is process_radar_cond(cube): it depends only on the cube, so evaluating it once instead of
    inside every net() call (reference :414-415) is bit-identical."""
    t_steps = karras_sigmas(num_steps, sigma_min, sigma_max, rho)
    x_next = latents.to(torch.float32) * t_steps[0]
    for i in range(num_steps):
        t_cur, t_next = t_steps[i], t_steps[i + 1]
        if S_churn > 0:
            gamma = min(S_churn / num_steps, math.sqrt(2) - 1) if S_min <= float(t_cur) <= S_max else 0.0
            t_hat = torch.as_tensor(t_cur + gamma * t_cur)
            x_hat = x_next + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * noises[i]
        else:
            t_hat, x_hat = t_cur, x_next
        denoised = edm_precond(sd, x_hat, t_hat, cond_tokens, heads=heads)
        d_cur = (x_hat - denoised) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            denoised = edm_precond(sd, x_next, t_next, cond_tokens, heads=heads)
            d_prime = (x_next - denoised) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
        if trace is not None:
            trace.append(x_next.clone())
        if stop_after is not None and i + 1 >= stop_after:  # tests: only the first steps of the full schedule
            break
    return x_next


# ----------------------------------------------------------------------------------------------
# radar encoder (models_radar_encoder.py) and tokenisation (models_radar_generation.py:363-407)
# ----------------------------------------------------------------------------------------------
def _gn_swish(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """GroupNorm(32, eps 1e-6, affine) then x*sigmoid(x): models_radar_encoder.py:5-12."""
    h = F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)
    return h * torch.sigmoid(h)


def _conv(sd: SD, name: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv3d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _resnet_block(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """models_radar_encoder.py:82-100 with temb=None, dropout 0."""
    h = _conv(sd, name + ".conv1", _gn_swish(sd, name + ".norm1", x))
    h = _conv(sd, name + ".conv2", _gn_swish(sd, name + ".norm2", h))
    if name + ".nin_shortcut.weight" in sd:
        x = _conv(sd, name + ".nin_shortcut", x, padding=0)
    return x + h


def _radar_attn_block(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """models_radar_encoder.py:112-135 — single head over all voxels, scale c^-0.5, GroupNorm WITHOUT swish."""
    h = F.group_norm(x, 32, sd[name + ".norm.weight"], sd[name + ".norm.bias"], eps=1e-6)
    q = _conv(sd, name + ".q", h, padding=0)
    k = _conv(sd, name + ".k", h, padding=0)
    v = _conv(sd, name + ".v", h, padding=0)
    b, c = q.shape[:2]
    qf = q.reshape(b, c, -1).transpose(1, 2)
    kf = k.reshape(b, c, -1)
    w = torch.bmm(qf, kf) * (int(c) ** -0.5)
    w = F.softmax(w, dim=2)
    vf = v.reshape(b, c, -1)
    o = torch.bmm(vf, w.transpose(1, 2)).reshape(x.shape)
    return x + _conv(sd, name + ".proj_out", o, padding=0)


def radar_encoder(sd: SD, x: torch.Tensor, prefix: str = "radar_enc.") -> torch.Tensor:
    """Encoder.forward, models_radar_encoder.py:216-241. x: [B, Cin, R, A, E] -> [B, z, R/16, A/16, E/16]."""
    h = _conv(sd, prefix + "conv_in", x)
    level = 0
    while f"{prefix}down.{level}.block.0.conv1.weight" in sd:
        blk = 0
        while f"{prefix}down.{level}.block.{blk}.conv1.weight" in sd:
            h = _resnet_block(sd, f"{prefix}down.{level}.block.{blk}", h)
            if f"{prefix}down.{level}.attn.{blk}.q.weight" in sd:
                h = _radar_attn_block(sd, f"{prefix}down.{level}.attn.{blk}", h)
            blk += 1
        if f"{prefix}down.{level}.downsample.conv.weight" in sd:
            # pad the HIGH side of each spatial dim by one, then 3x3x3 stride-2 conv without padding (:37-41)
            h = F.pad(h, (0, 1, 0, 1, 0, 1))
            h = _conv(sd, f"{prefix}down.{level}.downsample.conv", h, stride=2, padding=0)
        level += 1
    h = _resnet_block(sd, prefix + "mid.block_1", h)
    h = _radar_attn_block(sd, prefix + "mid.attn_1", h)
    h = _resnet_block(sd, prefix + "mid.block_2", h)
    h = _gn_swish(sd, prefix + "norm_out", h)
    return _conv(sd, prefix + "conv_out", h)


def process_radar_cond(sd: SD, cube: torch.Tensor, use_encoder: bool = True) -> torch.Tensor:
    """models_radar_generation.py:363-407. cube [B,R,A,E,2] -> tokens [B, r*a*e, C] (r-major, then a, then e)."""
    x = cube[..., 0:1]
    if use_encoder:
        x = radar_encoder(sd, x.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    tok = _lin(sd, "radar_token_project", x)
    r_emb, a_emb, e_emb = sd["radar_r_emb.weight"], sd["radar_a_emb.weight"], sd["radar_e_emb.weight"]
    B, r, a, e, C = tok.shape
    tok = tok + r_emb[:r][None, :, None, None, :] + a_emb[:a][None, None, :, None, :] + e_emb[:e][None, None, None, :, :]
    return tok.reshape(B, -1, C)


# ----------------------------------------------------------------------------------------------
# VecSet KL autoencoder (models_ae.py)
# ----------------------------------------------------------------------------------------------
def point_embed(sd: SD, p: torch.Tensor, prefix: str = "point_embed.") -> torch.Tensor:
    """models_ae.py:128-138 — [sin(p·basis) (24), cos(p·basis) (24), p (3)] -> Linear(51, dim)."""
    proj = torch.einsum("bnd,de->bne", p, sd[prefix + "basis"])
    feat = torch.cat([proj.sin(), proj.cos(), p], dim=2)
    return _lin(sd, prefix + "mlp", feat)


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"])


def _ae_attention(sd: SD, name: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """PreNorm(Attention): models_ae.py:41-49 + 84-105. `name` is the PreNorm module; context is LayerNormed
    only when the PreNorm owns a norm_context."""
    xn = _ln(sd, name + ".norm", x)
    if context is None:
        ctx = xn
    elif name + ".norm_context.weight" in sd:
        ctx = _ln(sd, name + ".norm_context", context)
    else:
        ctx = context
    q = F.linear(xn, sd[name + ".fn.to_q.weight"])
    k, v = F.linear(ctx, sd[name + ".fn.to_kv.weight"]).chunk(2, dim=-1)
    o = _heads_attention(q, k, v, heads)
    return _lin(sd, name + ".fn.to_out", o)


def _ae_ff(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    xn = _ln(sd, name + ".norm", x)
    return _geglu_ff(xn, sd[name + ".fn.net.0.weight"], sd[name + ".fn.net.0.bias"], sd[name + ".fn.net.2.weight"],
                     sd[name + ".fn.net.2.bias"])


def fps_indices(pc: torch.Tensor, m: int) -> torch.Tensor:
    """Farthest point sampling per cloud with start index 0 and lowest-index tie-break; returns int64 [B, m]
    indices INTO EACH CLOUD. Restates torch_cluster==1.6.3 fps (requirements.txt:156; call sites
    models_ae.py:243, 368) with random_start=False. Squared distance = (dx*dx + dy*dy) + dz*dz in fp32
    without FMA contraction (numpy elementwise ops are individually rounded)."""
    pts = pc.detach().cpu().numpy().astype(np.float32)
    B, N, _ = pts.shape
    out = np.zeros((B, m), dtype=np.int64)
    for b in range(B):
        p = pts[b]
        dist = np.full((N,), np.inf, dtype=np.float32)
        cur = 0
        for i in range(m):
            out[b, i] = cur
            d = p - p[cur]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            dist = np.minimum(dist, d2)
            cur = int(np.argmax(dist))  # first maximum = lowest index
    return torch.from_numpy(out)


def ae_encode_stats(sd: SD, pc: torch.Tensor, query_type: str, num_latents: int = 512):
    """KLAutoEncoder.encode up to the posterior parameters, models_ae.py:351-399. Returns (mean, logvar)
    with logvar already clamped to [-30, 20] (:145)."""
    B, N, _ = pc.shape
    pe = point_embed(sd, pc)
    if query_type == "point":
        idx = fps_indices(pc, num_latents)
        sampled = torch.gather(pc, 1, idx[..., None].expand(-1, -1, 3))
        x = point_embed(sd, sampled)
    elif query_type == "learnable":
        x = sd["latents.weight"][None].expand(B, -1, -1)
    elif query_type == "mix":
        dq = sd["d_latents.weight"][None].expand(B, -1, -1)
        dq = _ae_attention(sd, "mix_attn_layer", dq, pe, heads=8)  # context NOT normalised (no norm_context)
        x = _lin(sd, "query_proj", sd["s_latents.weight"][None] + dq)
    else:
        raise NotImplementedError(query_type)
    x = _ae_attention(sd, "cross_attend_blocks.0", x, pe, heads=1) + x
    x = _ae_ff(sd, "cross_attend_blocks.1", x) + x
    if "mean_fc.weight" not in sd:
        return x          # deterministic AutoEncoder.encode, models_ae.py:226-257: the latents are x itself
    mean = _lin(sd, "mean_fc", x)
    logvar = torch.clamp(_lin(sd, "logvar_fc", x), -30.0, 20.0)
    return mean, logvar


def ae_posterior(mean: torch.Tensor, logvar: torch.Tensor, noise: torch.Tensor):
    """DiagonalGaussianDistribution.sample / kl, models_ae.py:147-163 with the noise injected."""
    z = mean + torch.exp(0.5 * logvar) * noise
    kl = 0.5 * torch.mean(mean.pow(2) + torch.exp(logvar) - 1.0 - logvar, dim=[1, 2])
    return kl, z


def ae_depth(sd: SD) -> int:
    n = 0
    while f"layers.{n}.0.fn.to_q.weight" in sd:
        n += 1
    return n


def ae_latent_stack(sd: SD, z: torch.Tensor) -> torch.Tensor:
    """proj + 24 x (self-attn, FF) of KLAutoEncoder.decode, models_ae.py:410-414."""
    x = _lin(sd, "proj", z) if "proj.weight" in sd else z   # deterministic AutoEncoder.decode (:260-264) has no proj
    for n in range(ae_depth(sd)):
        x = _ae_attention(sd, f"layers.{n}.0", x, None, heads=8) + x
        x = _ae_ff(sd, f"layers.{n}.1", x) + x
    return x


def ae_query(sd: SD, x: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    """decoder cross attention + to_outputs, models_ae.py:417-424 (no residual, decoder_ff=False)."""
    qe = point_embed(sd, queries)
    lat = _ae_attention(sd, "decoder_cross_attn", qe, x, heads=1)
    return _lin(sd, "to_outputs", lat)


def ae_decode(sd: SD, z: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    return ae_query(sd, ae_latent_stack(sd, z), queries)


# ----------------------------------------------------------------------------------------------
# metrics (utils/utils.py:116-142) and helpers
# ----------------------------------------------------------------------------------------------
def chamfer_distance(pred: np.ndarray, gt: np.ndarray) -> float:
    """0.5*mean(NN dist pred->gt) + 0.5*mean(NN dist gt->pred), Euclidean (utils/utils.py:116-142)."""
    from scipy.spatial import cKDTree
    if len(pred) == 0 or len(gt) == 0:
        return float("inf")
    d1, _ = cKDTree(gt).query(pred)
    d2, _ = cKDTree(pred).query(gt)
    return 0.5 * float(np.mean(d1)) + 0.5 * float(np.mean(d2))


def occupancy_points(logits: np.ndarray, queries: np.ndarray, pc_range=None, norm_anisotropy: bool = True,
                     norm_isotropy: bool = False, view_cone: bool = False, threshold: float = 0.0) -> np.ndarray:
    """One frame: logits [Q], queries [Q, 3] float32 -> occupied points [P, 3] float32.
    engine_generation.py:283-289 (np.where(output > 0), gather, inverse_norm_points = utils/utils.py:50-76) and
    :313-315 (polar2cartesian = dataset_preprocessor/lidar.py:57-63), numpy float32 as in the reference."""
    ind = np.where(logits > threshold)[0]
    pts = queries[ind].astype(np.float32)
    if pc_range is not None:
        r = [float(v) for v in pc_range]
        off = [(r[3] + r[0]) / 2, (r[4] + r[1]) / 2, (r[5] + r[2]) / 2]
        sc = [(r[3] - r[0]) / 2, (r[4] - r[1]) / 2, (r[5] - r[2]) / 2]
        pred = np.zeros_like(pts)
        if norm_anisotropy:
            for a in range(3):
                pred[:, a] = pts[:, a] * sc[a] + off[a]
        if norm_isotropy:
            pred[:, :3] = pts[:, :3] * max(sc) + np.array(off)
        pts = pred
    if view_cone:
        r_, az, el = pts[:, 0], -np.deg2rad(pts[:, 1]), np.deg2rad(pts[:, 2])
        pts = np.stack([r_ * np.cos(el) * np.cos(az), r_ * np.cos(el) * np.sin(az), r_ * np.sin(el)], axis=1)
    return pts


def norm_constants(pc_range, norm_anisotropy: bool = True, norm_isotropy: bool = False):
    """(scale[3], offset[3]) python floats shared by norm_points / inverse_norm_points (utils/utils.py:60-65, 88-93)."""
    r = [float(v) for v in pc_range]
    off = [(r[3] + r[0]) / 2, (r[4] + r[1]) / 2, (r[5] + r[2]) / 2]
    sc = [(r[3] - r[0]) / 2, (r[4] - r[1]) / 2, (r[5] - r[2]) / 2]
    if norm_isotropy:
        sc = [max(sc)] * 3
    return sc, off


def norm_points(points: np.ndarray, pc_range, norm_anisotropy: bool = True, norm_isotropy: bool = False) -> np.ndarray:
    """utils/utils.py:78-104: (p - offset) / scale per axis in the array's own dtype (float32 here)."""
    sc, off = norm_constants(pc_range, norm_anisotropy, norm_isotropy)
    out = np.zeros_like(points)
    for a in range(3):
        out[:, a] = (points[:, a] - off[a]) / sc[a]
    return out


def aug_query_helper(helper_points: np.ndarray, aug_num: int, pc_range, voxel_size, aug_bias_scale: int,
                     sel: Optional[np.ndarray], scales: Optional[np.ndarray], u: Optional[np.ndarray]) -> np.ndarray:
    """datasets/utils/query_helper.py:3-43 with the three np.random draws injected (sel = choice(N, G),
    scales = choice(arange(aug_bias_scale) + 1, G), u = rand(G, 3)): float32 helper points + float64 bias, clipped to
    the range in float64, stored as float32."""
    N = helper_points.shape[0]
    out = np.zeros((aug_num, 3), np.float32)
    if N >= aug_num:
        out[:] = helper_points[:aug_num]
        return out
    vs = np.asarray(voxel_size, dtype=np.float64)
    bias = (u * 2 - 1) * (vs * scales[:, None])
    aug = helper_points[sel] + bias
    aug = np.clip(aug, np.asarray(pc_range[:3], dtype=np.float64), np.asarray(pc_range[3:], dtype=np.float64))
    out[:N] = helper_points
    out[N:] = aug
    return out


def process_radar_data(raw: np.ndarray, norm_intensity: bool, max_intensity: float, norm_dopp: bool, max_dopp: float,
                       upsample: bool, tgt_a: int, tgt_e: int) -> np.ndarray:
    """Coloradar_dataset.process_radar_data (datasets/aligned_coloradar/Coloradar_dataset.py:432-475) for one raw
    cube [R, A, E, C] float32 -> [R, A_up, E_up, 2] float32; the bilinear upsample is written out element by element
    the way ATen's CPU kernel evaluates it in the torch build of this image (align_corners=True; fp32 source index
    scale * dst; weights w_ij = w_a[i] * w_e[j] rounded to fp32; out = fma(x00, w00, fma(x01, w01, fma(x11, w11,
    x10 * w10))) — found by enumeration against F.interpolate, bit-exact on the fixture)."""
    raw = raw.astype(np.float32)
    R, A, E, _ = raw.shape
    out = np.zeros((R, A, E, 2), np.float32)
    if norm_intensity:
        out[..., 0] = np.clip(raw[..., 0], 0, max_intensity) / np.float32(max_intensity)
    out[..., 1] = raw[..., 1] * raw[..., -1]
    if norm_dopp:
        out[..., 1] = out[..., 1] / np.float32(max_dopp)
    if not upsample:
        return out

    def axis(n_in, n_out):
        f32 = np.float32
        if n_out == n_in:
            i0 = np.arange(n_out)
            return i0, i0, np.ones(n_out, f32), np.zeros(n_out, f32)
        scale = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)
        src = (scale * np.arange(n_out, dtype=f32)).astype(f32)
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        lam = np.clip((src - i0.astype(f32)).astype(f32), f32(0), f32(1))
        i1 = i0 + (i0 < n_in - 1)
        return i0, i1, (f32(1) - lam).astype(f32), lam

    a0, a1, wa0, wa1 = axis(A, tgt_a)
    e0, e1, we0, we1 = axis(E, tgt_e)
    def fma(a, b, c):   # fp32 fused multiply-add through fp64 (the product is exact in fp64)
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    res = np.zeros((R, tgt_a, tgt_e, 2), np.float32)
    w00 = wa0[:, None] * we0[None, :]
    w01 = wa0[:, None] * we1[None, :]
    w10 = wa1[:, None] * we0[None, :]
    w11 = wa1[:, None] * we1[None, :]
    for ch in range(2):
        x = out[..., ch]
        x00, x01 = x[:, a0][:, :, e0], x[:, a0][:, :, e1]
        x10, x11 = x[:, a1][:, :, e0], x[:, a1][:, :, e1]
        bc = np.broadcast_to
        res[..., ch] = fma(x00, bc(w00, x00.shape), fma(x01, bc(w01, x00.shape), fma(x11, bc(w11, x00.shape), x10 * w10)))
    return res


# ----------------------------------------------------------------------------------------------
# training / AE-evaluation loop helpers
# ----------------------------------------------------------------------------------------------
def update_ema(target_params, source_params, rate: float = 0.99) -> None:
    """engine_generation.py:29-39, in numpy with the fp32 roundings spelled out: t1 = rn32(targ * rn32(rate));
    targ = rn32(t1 + src * rn32(1 - rate)) with the product unrounded (ATen contracts ``a + alpha * b`` into one fma —
    the fp64 product of two fp32 values is exact, so one fp64 add followed by the fp32 rounding is that fma up to
    double rounding). In place on numpy arrays or torch CPU tensors."""
    r32, a32 = np.float32(rate), np.float32(1 - rate)
    for targ, src in zip(target_params, source_params):
        t = targ.detach().numpy() if isinstance(targ, torch.Tensor) else targ
        s = src.detach().numpy() if isinstance(src, torch.Tensor) else src
        t1 = (t * r32).astype(np.float32)
        t[...] = (t1.astype(np.float64) + s.astype(np.float64) * np.float64(a32)).astype(np.float32)


def occupancy_iou(logits: torch.Tensor, labels: torch.Tensor, threshold: float = 0.0):
    """engine_generation.py:376-385 (cache_latents; engine_ae.py's evaluate has the same lines), per frame (the
    reference takes the batch mean of each): logits, labels [B, Q] -> (accuracy [B], iou [B])."""
    pred = torch.zeros_like(logits)
    pred[logits >= threshold] = 1
    accuracy = (pred == labels).float().sum(dim=1) / labels.shape[1]
    intersection = (pred * labels).sum(dim=1)
    union = (pred + labels).gt(0).sum(dim=1)
    iou = intersection * 1.0 / union + 1e-5
    return accuracy, iou


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def fps_indices_c(pc: torch.Tensor, m: int) -> torch.Tensor:
    """Same as fps_indices through the plain-C restatement oracle/fps_oracle.c (fast enough for 64 clouds of 10000
    points); oracle/build_oracle.py compiles it with gcc -ffp-contract=off."""
    import ctypes
    from oracle import build_oracle
    lib = ctypes.CDLL(str(build_oracle.build()))
    pts = np.ascontiguousarray(pc.detach().cpu().numpy().astype(np.float32))
    B, N, _ = pts.shape
    out = np.zeros((B, m), dtype=np.int64)
    rc = lib.fps_oracle(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(m),
                        out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return torch.from_numpy(out)
