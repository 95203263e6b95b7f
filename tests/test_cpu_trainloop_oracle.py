"""CPU: oracle restatements of update_ema (engine_generation.py:29-39) and the accuracy / IoU statements of
cache_latents (engine_generation.py:376-385) against the fixture written from the UNMODIFIED reference code
(tests/golden/make_golden_trainloop.py), bit-exact; the cache writers of rald_b200.cache_io (host file I/O, SURVEY.md
§8f row 4) read back; update_ema / occupancy_iou have no CPU path."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import rald_oracle as orc
from rald_b200 import _lib, cache_io, postproc, train


def _g():
    return np.load(os.path.join(GOLDEN, "trainloop.npz"))


def _ema_lists(g):
    n = sum(1 for k in g.files if k.startswith("ema_target_"))
    return ([g[f"ema_target_{i}"].copy() for i in range(n)], [g[f"ema_source_{i}"] for i in range(n)],
            [g[f"ema_after3_{i}"] for i in range(n)], [g[f"ema_default_{i}"] for i in range(n)])


def test_update_ema_oracle_bit_exact():
    g = _g()
    targ, src, after3, default = _ema_lists(g)
    t = [a.copy() for a in targ]
    for _ in range(3):
        orc.update_ema(t, src, rate=float(g["ema_rate"]))
    assert all(np.array_equal(a, b) for a, b in zip(t, after3))
    t = [a.copy() for a in targ]
    orc.update_ema(t, src)
    assert all(np.array_equal(a, b) for a, b in zip(t, default))


def test_update_ema_oracle_matches_aten_cpu():
    """The fma form of the restatement is ATen's CPU arithmetic on a million elements (an unfused add differs in
    hundreds of them)."""
    gen = torch.Generator().manual_seed(0)
    t, s = torch.randn(1000003, generator=gen), torch.randn(1000003, generator=gen)
    ref = t.clone()
    ref.mul_(0.9999).add_(s, alpha=1 - 0.9999)
    mine = [t.clone().numpy()]
    orc.update_ema(mine, [s.numpy()], rate=0.9999)
    assert np.array_equal(mine[0], ref.numpy())


def test_occupancy_iou_oracle_bit_exact():
    g = _g()
    acc, iou = orc.occupancy_iou(torch.from_numpy(g["iou_logits"]), torch.from_numpy(g["iou_labels"]).float(), 0.0)
    assert np.array_equal(acc.numpy(), g["iou_accuracy"])
    assert np.array_equal(iou.numpy(), g["iou_iou"], equal_nan=True)
    assert np.isnan(g["iou_iou"][4])        # empty union: 0 / 0 + 1e-5, as the reference computes it


def test_no_cpu_path():
    with pytest.raises(_lib.RaldError):
        train.update_ema([torch.zeros(4)], [torch.ones(4)])
    with pytest.raises(_lib.RaldError):
        postproc.occupancy_iou(torch.zeros(1, 8), torch.zeros(1, 8))


def test_ply_round_trip(tmp_path):
    pts = np.random.RandomState(0).randn(1234, 3).astype(np.float32)
    p = tmp_path / "a.ply"
    cache_io.write_ply(p, torch.from_numpy(pts))
    raw = p.read_bytes()
    head = raw[:raw.index(b"end_header\n") + len(b"end_header\n")].decode()
    assert head.splitlines()[:3] == ["ply", "format binary_little_endian 1.0", "comment Created by Open3D"]
    assert "element vertex 1234" in head and head.count("property double") == 3
    assert len(raw) == len(head) + 1234 * 24
    back = cache_io.read_ply(p)
    assert back.dtype == np.float64 and np.array_equal(back, pts.astype(np.float64))
    cache_io.write_ply(p, np.zeros((0, 3)))
    assert cache_io.read_ply(p).shape == (0, 3)


def test_store_paths_and_payloads(tmp_path):
    """Paths as engine_generation.py:209-222, 324-338, 398-409 build them."""
    B = 2
    radar = [f"/data/seqA/radar/cube/{i:06d}.bin" for i in range(B)]
    lidar = [f"/data/seqA/lidar/{i:06d}.bin" for i in range(B)]
    pts = torch.arange(B * 5 * 3, dtype=torch.float32).reshape(B, 5, 3)
    cnt = torch.tensor([3, 0], dtype=torch.int32)
    files = cache_io.store_point_clouds(pts, cnt, radar, tmp_path, "exp", "pc")
    assert [str(f.relative_to(tmp_path)) for f in files] == ["exp/seqA/pc/000000.ply", "exp/seqA/pc/000001.ply"]
    assert np.array_equal(cache_io.read_ply(files[0]), pts[0, :3].double().numpy())
    assert cache_io.read_ply(files[1]).shape == (0, 3)

    z = torch.randn(B, 512, 32)
    files = cache_io.store_latent_tokens(z, lidar, radar, tmp_path, "exp")
    assert [str(f.relative_to(tmp_path)) for f in files] == ["exp/seqA/latent_tokens/000000.pt",
                                                             "exp/seqA/latent_tokens/000001.pt"]
    assert torch.equal(torch.load(files[1]), z)                      # the reference stores the whole batch per file
    files = cache_io.store_latent_tokens(z, lidar, radar, tmp_path, "exp", per_frame=True)
    assert torch.equal(torch.load(files[1]), z[1:2])

    files = cache_io.cache_latent_npz(z, lidar, tmp_path / "cache")
    assert [str(f.relative_to(tmp_path)) for f in files] == ["cache/seqA/000000.bin.npz", "cache/seqA/000001.bin.npz"]
    assert np.array_equal(np.load(files[0])["res_tokens"], z[0].numpy())
