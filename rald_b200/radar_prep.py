"""Radar-cube input preparation on the device: the reference's dataset does this per frame in
numpy + a CPU F.interpolate inside the dataloader worker and ships the 2 MB upsampled cube
(datasets/aligned_coloradar/Coloradar_dataset.py:432-475, process_radar_data). Here the raw [R, A, E, 3] cube
(24 KB at the default 128 x 8 x 2) is what crosses PCIe; the [R, A_up, E_up, 2] conditioning cube that
EDMPrecond.sample / process_radar_cond take is produced by one kernel."""
from __future__ import annotations

import torch

from . import _lib


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


@torch.no_grad()
def process_radar_data(radar_cube: torch.Tensor, radar_cfg, early_return: bool = False, channels_out: int = 2
                       ) -> torch.Tensor:
    """radar_cube [B, R, A, E, C] or [R, A, E, C] fp32 on the device (C >= 2: intensity, doppler, ..., valid mask),
    radar_cfg = the ``dataset.radar`` node of the reference's config (norm_intensity, max_intensity, norm_dopp,
    max_dopp, upsample, input_r_dim, tgt_{r,a,e}_dim). Returns [B, R, A_up, E_up, channels_out] fp32 (batch axis kept
    only if given). ``early_return`` mirrors the reference flag: no doppler normalisation and no upsampling.
    channels_out=1 emits the intensity channel only — what the radar encoder reads."""
    if radar_cube.device.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    squeeze = radar_cube.dim() == 4
    x = (radar_cube[None] if squeeze else radar_cube).contiguous().float()
    B, R, A, E, C = x.shape
    norm_i = bool(_cfg_get(radar_cfg, "norm_intensity", False))
    norm_d = bool(_cfg_get(radar_cfg, "norm_dopp", False)) and not early_return
    up = bool(_cfg_get(radar_cfg, "upsample", False)) and not early_return
    a_up, e_up = A, E
    if up:
        tgt_r = _cfg_get(radar_cfg, "tgt_r_dim", R)
        if int(tgt_r) != R:
            raise AssertionError(f"Input radar cube r_dim {R} and target r_dim {tgt_r} do not match")
        a_up, e_up = int(_cfg_get(radar_cfg, "tgt_a_dim")), int(_cfg_get(radar_cfg, "tgt_e_dim"))
    out = torch.empty(B, R, a_up, e_up, int(channels_out), device=x.device, dtype=torch.float32)
    _lib.call("rald_radar_cube_prep", x.data_ptr(), B, R, A, E, C, a_up, e_up, int(channels_out), 1 if norm_i else 0,
              float(_cfg_get(radar_cfg, "max_intensity", 1.0) or 1.0), 1 if norm_d else 0,
              float(_cfg_get(radar_cfg, "max_dopp", 1.0) or 1.0), out.data_ptr(), _lib.cur_stream())
    return out[0] if squeeze else out
