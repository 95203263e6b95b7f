"""GPU parity of KLAutoEncoder.encode (point features, FPS, long-context attention through tcgen05 GEMMs, posterior)
and of the device-side occupancy post-processing, through the C ABI, against the oracle and the reference fixtures.

Bars: FPS indices and compaction indices bit-exact (integer work); mean / logvar within 1e-2 relative L2 (bf16
operands, fp32 accumulate); inverse-normalised coordinates bit-exact, polar->cartesian within 2 ulp-level 1e-6."""
import numpy as np
import pytest
import torch

from helpers import build_ae, rel_l2
from oracle import rald_oracle as orc
from rald_b200 import postproc, synth

pytestmark = pytest.mark.gpu
PC_RANGE = [0, -90, -20, 15.8, 90, 20]


@pytest.fixture(scope="module")
def ae_point():
    return build_ae("kl_d512_m512_l32", device="cuda")


@pytest.fixture(scope="module")
def ae_mix():
    return build_ae("kl_d512_m512_l32_mix", device="cuda")


def test_fps_matches_fixture_bit_exact(ae_point, golden):
    pc = synth.frustum_points(1, 10000, seed=1024)
    idx = ae_point._runtime().fps(pc.cuda(), 512).cpu()
    assert torch.equal(idx, golden("ae")["point_fps_idx"])


@pytest.mark.parametrize("kind", ["uniform", "frustum", "lattice", "duplicates"])
def test_fps_matches_c_oracle_bit_exact(ae_point, kind):
    if kind == "uniform":
        pc = synth.lidar_points(4, 10000, seed=5)
    elif kind == "frustum":
        pc = synth.frustum_points(3, 9999, seed=6)        # N not a multiple of the block size
    elif kind == "lattice":
        g = torch.stack(torch.meshgrid(*[torch.arange(16.0)] * 3, indexing="ij"), -1).reshape(1, 4096, 3) / 8 - 1
        pc = g.repeat(2, 1, 1)                            # thousands of exact ties
    else:
        pc = synth.lidar_points(1, 2048, seed=7)
        pc[:, 1000:] = pc[:, :1048].clone()                       # duplicated points
    m = 512 if pc.shape[1] >= 2048 else 64
    got = ae_point._runtime().fps(pc.cuda(), m).cpu()
    assert torch.equal(got, orc.fps_indices_c(pc, m))
    assert torch.equal(got[:1], orc.fps_indices(pc[:1], m))   # numpy restatement agrees with the C one


@pytest.mark.parametrize("n,m", [(16384, 512), (4097, 33), (4096, 64), (40, 40), (32, 1)])
def test_fps_size_limits_match_c_oracle(ae_point, n, m):
    """Largest cloud one CTA holds, sizes around the 48 KB shared-memory opt-in, every point sampled, a single pick."""
    pc = synth.lidar_points(2, n, seed=n + m)
    got = ae_point._runtime().fps(pc.cuda(), m).cpu()
    assert torch.equal(got, orc.fps_indices_c(pc, m))
    if m == n:
        assert sorted(got[0].tolist()) == list(range(n))
    with pytest.raises(Exception):
        ae_point._runtime().fps(synth.lidar_points(1, 16385, seed=1).cuda(), 8)


def test_fps_full_size_batch_properties(ae_point):
    """BASELINE-size batch (64 clouds x 10000 points): first pick 0, no repeats, greedy max-min property on a sample."""
    pc = synth.lidar_points(64, 10000, seed=11)
    idx = ae_point._runtime().fps(pc.cuda(), 512).cpu()
    assert torch.all(idx[:, 0] == 0)
    for b in (0, 31, 63):
        assert len(set(idx[b].tolist())) == 512
        p = pc[b].numpy()
        d = np.full(10000, np.inf, dtype=np.float32)
        for i in range(511):
            diff = p - p[int(idx[b, i])]
            d = np.minimum(d, (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2])
            assert d[int(idx[b, i + 1])] == d.max()


@pytest.mark.parametrize("qtype,name,pts", [("mix", "kl_d512_m512_l32_mix", "uniform"),
                                            ("point", "kl_d512_m512_l32", "frustum")])
def test_encode_stats_match_reference(ae_mix, ae_point, golden, qtype, name, pts):
    g = golden("ae")
    ae = ae_mix if qtype == "mix" else ae_point
    pc = synth.lidar_points(1, 10000, seed=1024) if pts == "uniform" else synth.frustum_points(1, 10000, seed=1024)
    mean, logvar = ae.encode_stats(pc.cuda())
    e_m, e_l = rel_l2(mean, g[f"{qtype}_mean"]), rel_l2(logvar, g[f"{qtype}_logvar"])
    print(f"[{qtype}] mean rel-L2 {e_m:.3e}  logvar rel-L2 {e_l:.3e}")
    assert e_m < 1e-2 and e_l < 1e-2
    # encode(): noise from the global CPU generator, as the reference draws it (models_ae.py:153)
    torch.manual_seed(7)
    kl, z = ae.encode(pc.cuda())
    torch.manual_seed(7)
    kl_ref, z_ref = orc.ae_posterior(g[f"{qtype}_mean"], g[f"{qtype}_logvar"], torch.randn(1, 512, 32))
    assert rel_l2(z, z_ref) < 1e-2
    assert rel_l2(kl, g[f"{qtype}_kl"]) < 1e-2


def test_encode_batch_equals_single_frames(ae_mix):
    pc = synth.lidar_points(2, 10000, seed=21).cuda()
    m2, l2 = ae_mix.encode_stats(pc)
    m1, l1 = ae_mix.encode_stats(pc[1:2].contiguous())
    assert torch.equal(m2[1], m1[0]) and torch.equal(l2[1], l1[0])


def test_forward_returns_logits_and_kl(ae_mix):
    pc = synth.lidar_points(1, 10000, seed=3).cuda()
    q = synth.query_points(1, 777).cuda()
    out = ae_mix(pc, q)
    assert out["logits"].shape == (1, 777) and out["kl"].shape == (1,)
    assert torch.isfinite(out["logits"]).all()


@pytest.mark.parametrize("B,Q", [(1, 1000), (3, 4096), (2, 500000)])
def test_occupancy_compaction_is_np_where(B, Q):
    g = torch.Generator("cpu").manual_seed(B * 7 + Q)
    logits = torch.randn(B, Q, generator=g) - 1.2       # ~11 % occupied
    q = synth.query_points(B, Q, seed=Q)
    pts, cnt, idx = postproc.occupied_points(logits.cuda(), q.cuda(), 0.0, PC_RANGE, True, False, True,
                                             return_index=True)
    pts, cnt, idx = pts.cpu().numpy(), cnt.cpu().numpy(), idx.cpu().numpy()
    for b in range(B):
        ref_idx = np.where(logits[b].numpy() > 0)[0]
        assert cnt[b] == len(ref_idx)
        assert np.array_equal(idx[b, :cnt[b]], ref_idx)
        ref = orc.occupancy_points(logits[b].numpy(), q[b].numpy(), PC_RANGE, True, False, True)
        assert np.allclose(pts[b, :cnt[b]], ref, rtol=0, atol=2e-6 * 16)
    # inverse normalisation alone is bit-exact
    pts2, cnt2, _ = postproc.occupied_points(logits.cuda(), q.cuda(), 0.0, PC_RANGE, True, False, False)
    ref = orc.occupancy_points(logits[0].numpy(), q[0].numpy(), PC_RANGE, True, False, False)
    assert np.array_equal(pts2[0, :int(cnt2[0])].cpu().numpy(), ref)


def test_occupancy_edge_cases():
    q = synth.query_points(1, 300).cuda()
    none = postproc.occupied_points(torch.full((1, 300), -1.0).cuda(), q)
    assert int(none[1][0]) == 0
    full = postproc.occupied_points(torch.full((1, 300), 1.0).cuda(), q, capacity=100)   # truncated by capacity
    assert int(full[1][0]) == 300
    assert torch.equal(full[0][0], q[0, :100])
    lists = postproc.to_list(*postproc.occupied_points(torch.tensor([[1.0, -1.0, 2.0]]).cuda(), q[:, :3])[:2])
    assert lists[0].shape == (2, 3)
