"""Final gather of generated point clouds across the frame-sharded ranks (one process per GPU, NCCL over
NVLink/NVSwitch). Frames are independent, so this is the ONLY collective on the generation path: an all_gather of
the per-frame point counts followed by an all_gather of the padded point buffers (<= ~1 MB per frame). The reference
has no such step — each rank writes its own .ply files (engine_generation.py:324-338) — it is what lets one caller
see the whole batch. Works on any torch.distributed backend (the CPU tests use gloo)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_frames(total_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of frames owned by `rank`; earlier ranks take the remainder. Seeds must be the
    GLOBAL frame indices (EDMPrecond.sample(batch_seeds=...)), otherwise every rank would restart at seed 0
    (models_radar_generation.py:440-441)."""
    base, rem = divmod(total_frames, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_point_clouds(points: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """points [F, cap, 3] fp32 and counts [F] int32 of this rank's frames -> (points [sum F, cap_max, 3],
    counts [sum F]) on every rank, frames in global (rank-major) order. Ranks may hold different F and cap."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return points, counts
    world = dist.get_world_size(group)
    dev = points.device
    meta = torch.tensor([points.shape[0], points.shape[1]], device=dev, dtype=torch.int64)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    frames = [int(m[0]) for m in metas]
    cap = max(int(m[1]) for m in metas)
    fmax = max(frames)
    pad_pts = torch.zeros(fmax, cap, 3, device=dev, dtype=points.dtype)
    pad_pts[:points.shape[0], :points.shape[1]] = points
    pad_cnt = torch.zeros(fmax, device=dev, dtype=counts.dtype)
    pad_cnt[:counts.shape[0]] = counts
    all_pts = [torch.empty_like(pad_pts) for _ in range(world)]
    all_cnt = [torch.empty_like(pad_cnt) for _ in range(world)]
    dist.all_gather(all_cnt, pad_cnt, group=group)
    dist.all_gather(all_pts, pad_pts, group=group)
    pts = torch.cat([p[:f] for p, f in zip(all_pts, frames)])
    cnt = torch.cat([c[:f] for c, f in zip(all_cnt, frames)])
    return pts, cnt


def gather_latents(latents: torch.Tensor, group=None) -> torch.Tensor:
    """[F, M, C] per rank -> [sum F, M, C] on every rank (64 KB per frame)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return latents
    world = dist.get_world_size(group)
    n = torch.tensor([latents.shape[0]], device=latents.device, dtype=torch.int64)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    fmax = max(int(v) for v in ns)
    pad = torch.zeros(fmax, *latents.shape[1:], device=latents.device, dtype=latents.dtype)
    pad[:latents.shape[0]] = latents
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:int(k)] for o, k in zip(outs, ns)])
