"""Device-side post-processing of decoded occupancy logits — what the reference's ``evaluate`` does in numpy on the
host after copying every logit back (engine_generation.py:283-289, 313-315): threshold, gather the occupied query
points, inverse normalisation (utils/utils.py:50-76) and, in view-cone mode, polar -> cartesian
(dataset_preprocessor/lidar.py:57-63). Here only the occupied points leave the GPU."""
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
