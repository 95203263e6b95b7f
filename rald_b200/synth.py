"""Synthetic inputs of the reference's shapes. Everything is drawn on the CPU from seeded
generators so the oracle (CPU) and the CUDA path see identical tensors; no dataset is needed."""
from __future__ import annotations

import torch

from .config import RADAR_CUBE_SHAPE


def _gen(seed: int) -> torch.Generator:
    return torch.Generator("cpu").manual_seed(int(seed))


def radar_cube(batch: int, seed: int = 1024, sparse: bool = False) -> torch.Tensor:
    """[B, 128, 64, 32, 2] fp32. Channel 0 in [0, 1] like clip(dB, 0, 45)/45 (Coloradar_dataset.py:447-451);
    `sparse` keeps ~2 % strong returns over a low noise floor, closer to real range-azimuth-elevation cubes."""
    g = _gen(seed)
    x = torch.rand(batch, *RADAR_CUBE_SHAPE, generator=g)
    if sparse:
        keep = torch.rand(batch, *RADAR_CUBE_SHAPE[:3], generator=g) < 0.02
        x[..., 0] = torch.where(keep, 0.5 + 0.5 * x[..., 0], 0.05 * x[..., 0])
    return x


def lidar_points(batch: int, n: int = 10000, seed: int = 1024) -> torch.Tensor:
    """[B, N, 3] uniform in [-1, 1]^3 (normalised polar frustum coordinates, Coloradar_dataset.py:365-379)."""
    return 2.0 * torch.rand(batch, n, 3, generator=_gen(seed)) - 1.0


def frustum_points(batch: int, n: int = 10000, seed: int = 1024) -> torch.Tensor:
    """[B, N, 3] points on a few planar/cylindrical surfaces quantised to 1/2048 — exercises FPS ties."""
    g = _gen(seed)
    u = torch.rand(batch, n, 3, generator=g)
    which = (u[..., 2] * 4).floor()
    x = 2 * u[..., 0] - 1
    y = 2 * u[..., 1] - 1
    z = torch.where(which < 2, (which - 0.5) * 0.8 + 0 * x, 0.3 * torch.sin(3 * x) * torch.cos(2 * y))
    pts = torch.stack([x, y, z], dim=-1)
    return torch.round(pts * 2048) / 2048


def query_points(batch: int, q: int, seed: int = 4242) -> torch.Tensor:
    """[B, Q, 3] uniform in [-1, 1]^3 (utils/utils.py:171-175 draws uniform query grids)."""
    return 2.0 * torch.rand(batch, q, 3, generator=_gen(seed)) - 1.0


def unit_latents(seeds, n_latents: int = 512, channels: int = 32) -> torch.Tensor:
    """Per-frame unit normal latents from torch.Generator('cpu').manual_seed(seed): what
    StackedRandomGenerator('cpu', seeds).randn gives (models_radar_generation.py:297-304)."""
    return torch.stack([torch.randn(n_latents, channels, generator=_gen(int(s) % (1 << 32))) for s in seeds])


def posterior_noise(batch: int, m: int = 512, c: int = 32, seed: int = 7) -> torch.Tensor:
    return torch.randn(batch, m, c, generator=_gen(seed))
