"""Device-side post-processing of decoded occupancy logits — what the reference's ``evaluate`` does in numpy on the
host after copying every logit back (engine_generation.py:283-289, 313-315): threshold, gather the occupied query
points, inverse normalisation (utils/utils.py:50-76) and, in view-cone mode, polar -> cartesian
(dataset_preprocessor/lidar.py:57-63). Here only the occupied points leave the GPU.

Also: the ``refine_query`` second-pass query set (engine_generation.py:291-297,
datasets/utils/query_helper.py:3-43) built on the device from the first pass's occupied points, and the Chamfer
metric (utils/utils.py:116-142) as a brute-force nearest-neighbour kernel instead of per-point cKDTree queries."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def inverse_norm_constants(pc_range: Sequence[float], norm_anisotropy: bool = True, norm_isotropy: bool = False
                           ) -> np.ndarray:
    """{sx, sy, sz, ox, oy, oz} of inverse_norm_points as fp32 (isotropic normalisation wins when both flags are set,
    as in the reference where its assignment comes last)."""
    r = [float(v) for v in pc_range]
    off = [(r[3] + r[0]) / 2, (r[4] + r[1]) / 2, (r[5] + r[2]) / 2]
    sc = [(r[3] - r[0]) / 2, (r[4] - r[1]) / 2, (r[5] - r[2]) / 2]
    if norm_isotropy:
        sc = [max(sc)] * 3
    elif not norm_anisotropy:
        raise ValueError("inverse_norm_points returns zeros when neither normalisation flag is set")
    return np.asarray(sc + off, dtype=np.float32)


@torch.no_grad()
def occupied_points(logits: torch.Tensor, queries: torch.Tensor, threshold: float = 0.0,
                    pc_range: Optional[Sequence[float]] = None, norm_anisotropy: bool = True,
                    norm_isotropy: bool = False, view_cone: bool = False, capacity: Optional[int] = None,
                    return_index: bool = False) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """logits [B, Q] (or [B, Q, 1]), queries [B, Q, 3] on the device -> (points [B, cap, 3] fp32, counts [B] int32,
    index [B, cap] int32 or None). Row b holds its counts[b] occupied points in query order; rows are not cleared
    beyond that. capacity defaults to Q (no truncation possible)."""
    if logits.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    if logits.dim() == 3:
        logits = logits.squeeze(-1)
    B, Q = logits.shape
    logits = logits.contiguous().float()
    queries = queries.contiguous().float()
    cap = int(capacity) if capacity is not None else Q
    dev = logits.device
    points = torch.empty(B, cap, 3, device=dev, dtype=torch.float32)
    index = torch.empty(B, cap, device=dev, dtype=torch.int32) if return_index else None
    counts = torch.empty(B, device=dev, dtype=torch.int32)
    ws = torch.empty(int(_lib.lib().rald_occupancy_ws_elems(B, Q)), device=dev, dtype=torch.int32)
    so = inverse_norm_constants(pc_range, norm_anisotropy, norm_isotropy) if pc_range is not None else None
    _lib.call("rald_occupancy_compact", logits.data_ptr(), queries.data_ptr(), B, Q, float(threshold),
              so.ctypes.data if so is not None else 0, 1 if view_cone else 0, cap, points.data_ptr(), _lib.ptr(index),
              counts.data_ptr(), ws.data_ptr(), _lib.cur_stream())
    return points, counts, index


def to_list(points: torch.Tensor, counts: torch.Tensor) -> List[np.ndarray]:
    """Per-frame numpy clouds [P_i, 3] (one device->host copy of the occupied points only)."""
    n = counts.cpu().tolist()
    cap = points.shape[1]
    pts = points[:, :max(1, min(cap, max(n) if n else 0))].cpu().numpy()
    return [pts[i, :min(k, cap)].copy() for i, k in enumerate(n)]


def polar_to_cartesian(points: torch.Tensor) -> torch.Tensor:
    """dataset_preprocessor/lidar.py:57-63 on a device tensor [..., 3] = (r, azimuth deg, elevation deg); used for
    the ground-truth side of the metric (the predicted side is converted inside rald_occupancy_compact)."""
    r, az, el = points[..., 0], -torch.deg2rad(points[..., 1]), torch.deg2rad(points[..., 2])
    return torch.stack([r * torch.cos(el) * torch.cos(az), r * torch.cos(el) * torch.sin(az), r * torch.sin(el)], -1)


def draw_refine_randoms(n_points: int, aug_num: int, aug_scale: int, rng=np.random):
    """The three draws of aug_query_helper in the reference's order and with its calls (query_helper.py:30-37), so that
    a run seeded like the reference (np.random.seed) refines with the same random numbers. Returns None when no rows
    are generated (n_points >= aug_num)."""
    gen = int(aug_num) - int(n_points)
    if gen <= 0:
        return None
    if n_points <= 0:
        raise ValueError("aug_query_helper needs at least one helper point (np.random.choice(0, ...) raises)")
    sel = rng.choice(int(n_points), size=gen, replace=True)
    scales = rng.choice(np.arange(aug_scale, step=1) + 1, size=gen)
    u = rng.rand(gen, 3)
    return sel.astype(np.int32), scales.astype(np.int32), np.ascontiguousarray(u, dtype=np.float64)


@torch.no_grad()
def refine_queries(points: torch.Tensor, count: torch.Tensor, aug_num: int, pc_range: Sequence[float],
                   voxel_size: Sequence[float], aug_scale: int = 2, norm_anisotropy: bool = True,
                   norm_isotropy: bool = False, rng="device", seed: int = 0) -> torch.Tensor:
    """Normalised second-pass queries [aug_num, 3] of ONE frame from its first-pass occupied points.

    points [cap, 3] fp32 (inverse-normalised polar points, a row of ``occupied_points(...)`` computed WITHOUT
    view_cone) and count (int32 device scalar / 1-element tensor) stay on the device. ``rng="numpy"`` reproduces the
    reference's np.random call sequence on the host (needs count on the host: one 4-byte read) and uploads the draws;
    ``rng="device"`` draws with Philox on the device (no host round trip)."""
    if points.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    points = points.contiguous().float()
    cap = points.shape[0]
    count = count.reshape(-1)[:1].to(torch.int32).contiguous()
    dev = points.device
    out = torch.empty(int(aug_num), 3, device=dev, dtype=torch.float32)
    so = inverse_norm_constants(pc_range, norm_anisotropy, norm_isotropy)
    voxel = np.asarray([float(v) for v in voxel_size], dtype=np.float64)
    rng_arr = np.asarray([float(v) for v in pc_range], dtype=np.float64)
    sel = scl = u = None
    if rng != "device":
        n = min(int(count.item()), cap)
        draws = draw_refine_randoms(n, aug_num, aug_scale, np.random if rng == "numpy" else rng)
        if draws is not None:
            pad = int(aug_num) - draws[0].shape[0]
            sel = torch.from_numpy(np.pad(draws[0], (0, pad))).to(dev)
            scl = torch.from_numpy(np.pad(draws[1], (0, pad), constant_values=1)).to(dev)
            u = torch.from_numpy(np.pad(draws[2], ((0, pad), (0, 0)))).to(dev)
        else:
            sel = torch.zeros(int(aug_num), device=dev, dtype=torch.int32)
            scl = torch.ones(int(aug_num), device=dev, dtype=torch.int32)
            u = torch.zeros(int(aug_num), 3, device=dev, dtype=torch.float64)
    _lib.call("rald_refine_queries", points.data_ptr(), count.data_ptr(), cap, int(aug_num), _lib.ptr(sel),
              _lib.ptr(scl), _lib.ptr(u), int(seed) & 0xFFFFFFFFFFFFFFFF, int(aug_scale), voxel.ctypes.data,
              rng_arr.ctypes.data, so.ctypes.data, out.data_ptr(), _lib.cur_stream())
    return out


@torch.no_grad()
def chamfer_distance(pred: torch.Tensor, pred_counts: torch.Tensor, gt: torch.Tensor,
                     gt_counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cal_metrics (utils/utils.py:116-142) for a batch on the device: pred [B, capP, 3] with pred_counts [B] valid
    rows per frame (exactly what ``occupied_points`` returns), gt [B, G, 3] (all rows valid unless gt_counts [B] is
    given). Returns float64 [B, 3] = (cd, mean NN distance pred->gt, mean NN distance gt->pred); inf for an empty
    side, as the reference returns np.inf for an empty prediction."""
    if pred.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    if pred.dim() == 2:
        pred, gt = pred[None], gt[None]
    pred = pred.contiguous().float()
    gt = gt.contiguous().float()
    B, cap_p, _ = pred.shape
    cap_g = gt.shape[1]
    if cap_p == 0 or cap_g == 0:
        return torch.full((B, 3), float("inf"), device=pred.device, dtype=torch.float64)
    pred_counts = pred_counts.reshape(-1).to(torch.int32).contiguous()
    if gt_counts is not None:
        gt_counts = gt_counts.reshape(-1).to(torch.int32).contiguous()
    out = torch.empty(B, 3, device=pred.device, dtype=torch.float64)
    ws = torch.empty(int(_lib.lib().rald_chamfer_ws_elems(B, cap_p, cap_g)), device=pred.device, dtype=torch.float64)
    _lib.call("rald_chamfer", pred.data_ptr(), pred_counts.data_ptr(), cap_p, gt.data_ptr(), _lib.ptr(gt_counts), cap_g,
              cap_g, B, out.data_ptr(), ws.data_ptr(), _lib.cur_stream())
    return out


@torch.no_grad()
def refine_pass(vae, latents: torch.Tensor, points: torch.Tensor, counts: torch.Tensor, aug_num: int,
                pc_range: Sequence[float], voxel_size: Sequence[float], aug_scale: int = 2, threshold: float = 0.0,
                norm_anisotropy: bool = True, norm_isotropy: bool = False, view_cone: bool = False,
                rng="device", seed: int = 0, capacity: Optional[int] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The ``refine_query`` branch of evaluate() (engine_generation.py:291-311) for a whole batch, on the device:
    per frame, aug_query_helper + norm_points on its first-pass occupied points (``points [B, cap, 3]`` / ``counts
    [B]`` from ``occupied_points(..., view_cone=False)``: aug_query_helper works in the polar frame) -> one batched
    ``vae.decode(latents, refined)`` -> threshold / inverse norm (/ polar2cartesian). The reference asserts batch 1 here
    and re-runs the 24-layer latent stack for the second decode; this decodes all frames in one call and the stack is
    reused from the first pass (runtime_ae caches it per latent tensor).
    Returns (refined points [B, cap2, 3], counts [B], refined queries [B, aug_num, 3])."""
    B = points.shape[0]
    queries = torch.stack([
        refine_queries(points[b], counts[b:b + 1], aug_num, pc_range, voxel_size, aug_scale, norm_anisotropy,
                       norm_isotropy, rng=rng, seed=seed + b) for b in range(B)])
    logits = vae.decode(latents, queries).squeeze(-1)
    pts, cnt, _ = occupied_points(logits, queries, threshold=threshold, pc_range=pc_range,
                                  norm_anisotropy=norm_anisotropy, norm_isotropy=norm_isotropy, view_cone=view_cone,
                                  capacity=capacity)
    return pts, cnt, queries


@torch.no_grad()
def occupancy_iou(logits: torch.Tensor, labels: torch.Tensor, threshold: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Accuracy and IoU of decoded occupancy against 0 / 1 labels, per frame (engine_generation.py:376-385 in
    ``cache_latents``; engine_ae.py's ``evaluate`` has the same lines): logits, labels [B, Q] on the device ->
    (accuracy [B], iou [B]) fp32 with ``pred = logits >= threshold``; the reference logs their batch means."""
    if logits.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    if logits.dim() == 3:
        logits = logits.squeeze(-1)
    if labels.shape != logits.shape:
        raise _lib.RaldError(f"occupancy_iou: logits {tuple(logits.shape)} vs labels {tuple(labels.shape)}")
    B, Q = logits.shape
    logits = logits.contiguous().float()
    labels = labels.to(logits.device).contiguous().float()
    out = torch.empty(B, 2, device=logits.device, dtype=torch.float32)
    ws = torch.empty(3 * B, device=logits.device, dtype=torch.int32)
    _lib.call("rald_occupancy_iou", logits.data_ptr(), labels.data_ptr(), B, Q, float(threshold), out.data_ptr(),
              ws.data_ptr(), _lib.cur_stream())
    return out[:, 0], out[:, 1]
