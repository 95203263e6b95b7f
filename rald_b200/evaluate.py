"""The per-batch body of the reference's ``evaluate`` (engine_generation.py:166-319) chained on the device:

    radar cube -> model.sample -> vae.decode(grid) -> threshold / inverse normalisation (/ polar2cartesian)
               [-> refine_query: aug_query_helper + norm_points -> second decode -> threshold ...]
               [-> Chamfer distance against the ground-truth surface]

Only the query grid (generated on the host with the reference's ``np.random.uniform`` calls, so a seeded run draws the
same grid) goes up and only occupied points / the metric come down. Nothing here is new arithmetic: every step is one
of the parity-tested pieces of ``rald_b200`` (models_*, postproc), in the reference's order."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import postproc


def generate_query_points(num_points: int, pc_range: Sequence[float], norm_anisotropy: bool = True,
                          norm_isotropy: bool = False, rng=np.random) -> np.ndarray:
    """utils/utils.py:148-176 (polar branch): three ``uniform(min, max, num_points)`` draws, x then y then z, stacked
    to float64 [num_points, 3] — [-1, 1]^3 for anisotropic normalisation, the scaled box for isotropic."""
    r = [float(v) for v in pc_range]
    sc = [(r[3] - r[0]) / 2, (r[4] - r[1]) / 2, (r[5] - r[2]) / 2]
    lo, hi = [-1.0] * 3, [1.0] * 3
    if norm_isotropy:
        m = max(sc)
        lo, hi = [-s / m for s in sc], [s / m for s in sc]
    elif not norm_anisotropy:
        raise ValueError("generate_query_points: neither normalisation flag is set (the reference raises NameError)")
    cols = [rng.uniform(lo[a], hi[a], int(num_points)) for a in range(3)]
    return np.stack(cols, axis=1)


@torch.no_grad()
def generate_point_clouds(model, vae, radar_cube: torch.Tensor, num_query_points: int, pc_range: Sequence[float],
                          norm_anisotropy: bool = True, norm_isotropy: bool = False, view_cone: bool = True,
                          threshold: float = 0.0, batch_seeds=None, refine: Optional[Dict] = None,
                          ground_truth: Optional[torch.Tensor] = None, grid: Optional[torch.Tensor] = None,
                          capacity: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """radar_cube [B, R, A, E, 2] on the device -> {"latents", "points" [B, cap, 3], "counts" [B], "cd" [B] (if
    ground_truth [B, G, 3], normalised polar coordinates like data_dict['lidar_points'], is given)}.

    refine = {"aug_num", "voxel_size", "scale", "rng" ("numpy" | "device"), "seed"} runs the ``refine_query`` second
    pass (any batch size; the reference asserts batch 1 there). ``grid`` overrides the generated query grid
    ([Q, 3] or [B, Q, 3], normalised)."""
    B = radar_cube.shape[0]
    dev = radar_cube.device
    z = model.sample(cond=radar_cube, batch_seeds=batch_seeds, cond_type="radar").to(torch.float32)
    if grid is None:
        grid_np = generate_query_points(num_query_points, pc_range, norm_anisotropy, norm_isotropy)
        grid = torch.from_numpy(grid_np.astype("float32")).to(dev, non_blocking=True)
    if grid.dim() == 2:
        grid = grid.unsqueeze(0).expand(B, -1, -1)
    grid = grid.contiguous().float()
    logits = vae.decode(z, grid).squeeze(-1)
    first_view_cone = view_cone and refine is None          # aug_query_helper works on the polar points
    pts, cnt, _ = postproc.occupied_points(logits, grid, threshold, pc_range, norm_anisotropy, norm_isotropy,
                                           first_view_cone, capacity=capacity)
    if refine is not None:
        pts, cnt, _ = postproc.refine_pass(vae, z, pts, cnt, int(refine["aug_num"]), pc_range, refine["voxel_size"],
                                           int(refine.get("scale", 2)), threshold, norm_anisotropy, norm_isotropy,
                                           view_cone, rng=refine.get("rng", "device"), seed=int(refine.get("seed", 0)),
                                           capacity=capacity)
    out = {"latents": z, "points": pts, "counts": cnt}
    if ground_truth is not None:
        so = torch.from_numpy(postproc.inverse_norm_constants(pc_range, norm_anisotropy, norm_isotropy)).to(dev)
        gt = ground_truth.to(dev).float() * so[:3] + so[3:]
        if view_cone:
            gt = postproc.polar_to_cartesian(gt)
        out["cd"] = postproc.chamfer_distance(pts, cnt, gt)[:, 0]
    return out
