"""Writers of the reference's eval / cache loops: the data formats on the far side of the hot
path. Host-side file I/O only; paths and payloads follow the reference so that its downstream scripts read them.

  * store_pc        engine_generation.py:324-338  — one ``.ply`` per frame (Open3D ``write_point_cloud`` defaults:
                    binary little-endian, double-precision x / y / z, comment line "Created by Open3D")
  * store_latent    engine_generation.py:209-222  — ``torch.save`` of the sampled tokens as ``<radar stem>.pt``
  * cache_latents   engine_generation.py:398-409  — ``np.savez(<frame>.npz, res_tokens=latents[idx])``
Open3D is not installed in this image, so the ``.ply`` layout is restated from its file format (rply writer), and
checked by reading the file back (tests/test_cpu_host_logic.py)."""
from __future__ import annotations

from pathlib import Path
from typing import Sequence, Union

import numpy as np
import torch

PathLike = Union[str, Path]


def write_ply(path: PathLike, points) -> None:
    """points [P, 3] (numpy or torch, any float dtype) -> binary little-endian PLY with double coordinates."""
    if isinstance(points, torch.Tensor):
        points = points.detach().cpu().numpy()
    pts = np.ascontiguousarray(np.asarray(points, dtype="<f8").reshape(-1, 3))
    header = ("ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
              f"element vertex {pts.shape[0]}\nproperty double x\nproperty double y\nproperty double z\nend_header\n")
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(pts.tobytes())


def read_ply(path: PathLike) -> np.ndarray:
    """Reads back what write_ply wrote (vertex element with x, y, z of one float type, binary little-endian)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        n, types = 0, []
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[:2] == ["format", "binary_little_endian"]:
                continue
            if tok[0] == "format":
                raise ValueError(f"read_ply: unsupported format line {line!r}")
            if tok[:2] == ["element", "vertex"]:
                n = int(tok[2])
            elif tok[0] == "property":
                types.append({"double": "<f8", "float": "<f4"}[tok[1]])
        dt = np.dtype([(f"c{i}", t) for i, t in enumerate(types)])
        raw = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return np.stack([raw[f"c{i}"].astype(np.float64) for i in range(3)], axis=1)


def store_point_clouds(points: torch.Tensor, counts: torch.Tensor, radar_paths: Sequence[PathLike], base_dir: PathLike,
                       exp_name: str, save_pc_dir_name: str) -> list:
    """engine_generation.py:324-338 for a whole batch: points [B, cap, 3] / counts [B] as returned by
    postproc.occupied_points; frame i goes to base_dir/exp_name/<sequence of radar_paths[i]>/save_pc_dir_name/<stem>.ply
    (sequence = the radar file's great-grandparent directory, as in the reference)."""
    pts = points.detach().cpu().numpy()
    cnt = counts.detach().cpu().numpy()
    written = []
    for i, rp in enumerate(radar_paths):
        rp = Path(rp)
        save_dir = Path(base_dir) / exp_name / rp.parent.parent.parent.name / save_pc_dir_name
        save_dir.mkdir(parents=True, exist_ok=True)
        out = save_dir / (rp.stem + ".ply")
        write_ply(out, pts[i, :int(cnt[i])])
        written.append(out)
    return written


def store_latent_tokens(sampled_tokens: torch.Tensor, lidar_paths: Sequence[PathLike], radar_paths: Sequence[PathLike],
                        base_dir: PathLike, exp_name: str, per_frame: bool = False) -> list:
    """engine_generation.py:209-222: base_dir/exp_name/<sequence of lidar_paths[i]>/latent_tokens/<radar stem>.pt.
    The reference saves the WHOLE batch tensor into every frame's file (it is run with batch size 1 there);
    per_frame=True stores ``sampled_tokens[i:i+1]`` instead, which is identical at batch size 1."""
    cpu = sampled_tokens.detach().cpu()
    written = []
    for i in range(cpu.shape[0]):
        seq = Path(lidar_paths[i]).parent.parent.name
        save_dir = Path(base_dir) / exp_name / seq / "latent_tokens"
        save_dir.mkdir(parents=True, exist_ok=True)
        out = save_dir / (Path(radar_paths[i]).stem + ".pt")
        torch.save(cpu[i:i + 1].clone() if per_frame else cpu, out)
        written.append(out)
    return written


def cache_latent_npz(res_latents: torch.Tensor, lidar_paths: Sequence[PathLike], cache_base_path: PathLike) -> list:
    """engine_generation.py:398-409: cache_base_path/<parts[-3] of the lidar path>/<file name>.npz with key
    ``res_tokens`` = that frame's latents [M, C] fp32."""
    lat = res_latents.detach().to(torch.float32).cpu().numpy()
    written = []
    for i, lp in enumerate(lidar_paths):
        lp = Path(lp)
        d = Path(cache_base_path) / lp.parts[-3]
        d.mkdir(parents=True, exist_ok=True)
        out = d / (lp.parts[-1] + ".npz")
        np.savez(out, res_tokens=lat[i])
        written.append(out)
    return written
