"""GPU parity of the eval-loop helpers through the C ABI against the fixture written from
the unmodified reference functions and against the oracle on seeded inputs.
Bars: refined queries and the prepared radar cube bit-exact (fp32 / fp64 arithmetic restated operation by operation);
Chamfer distance within 1e-6 relative (fp32 nearest-neighbour search, fp64 winner distance and sums)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import rald_oracle as orc
from rald_b200 import postproc, radar_prep

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _g():
    return np.load(os.path.join(GOLDEN, "evalpost.npz"))


def test_refine_queries_fixture_bit_exact():
    g = _g()
    helper = torch.from_numpy(g["refine_helper"]).to(DEV)
    n = helper.shape[0]
    pts = torch.cat([helper, torch.full((300, 3), 7.0, device=DEV)])       # rows beyond count must be ignored
    count = torch.tensor([n], device=DEV, dtype=torch.int32)
    np.random.seed(1234)
    q = postproc.refine_queries(pts, count, int(g["refine_aug_num"]), g["pc_range"].tolist(), g["voxel_size"].tolist(),
                                int(g["refine_scale"]), rng="numpy")
    assert np.array_equal(q.cpu().numpy(), g["refine_queries"])
    # truncation branch: more helper points than aug_num -> the first aug_num points, normalised
    q2 = postproc.refine_queries(pts, count, 1000, g["pc_range"].tolist(), g["voxel_size"].tolist(),
                                 int(g["refine_scale"]), rng="numpy")
    assert np.array_equal(q2.cpu().numpy(), orc.norm_points(g["refine_helper"][:1000], g["pc_range"].tolist()))


def test_refine_queries_device_rng_properties():
    g = _g()
    helper = torch.from_numpy(g["refine_helper"]).to(DEV)
    n, aug, scale = helper.shape[0], 200000, int(g["refine_scale"])
    count = torch.tensor([n], device=DEV, dtype=torch.int32)
    rng, vox = g["pc_range"], g["voxel_size"]
    q = postproc.refine_queries(helper, count, aug, rng.tolist(), vox.tolist(), scale, rng="device", seed=5)
    q_again = postproc.refine_queries(helper, count, aug, rng.tolist(), vox.tolist(), scale, rng="device", seed=5)
    q_other = postproc.refine_queries(helper, count, aug, rng.tolist(), vox.tolist(), scale, rng="device", seed=6)
    assert torch.equal(q, q_again) and not torch.equal(q, q_other)
    qn = q.cpu().numpy()
    assert np.array_equal(qn[:n], orc.norm_points(g["refine_helper"], rng.tolist()))
    assert np.abs(qn).max() <= 1.0
    # every generated point lies within aug_scale voxels of SOME helper point (checked against its nearest one)
    sc, off = orc.norm_constants(rng.tolist())
    gen = qn[n:] * np.asarray(sc, np.float32) + np.asarray(off, np.float32)
    from scipy.spatial import cKDTree
    d, j = cKDTree(g["refine_helper"] / vox).query(gen / vox, p=np.inf)
    assert d.max() <= scale + 1e-3
    # the bias is spread over the whole allowed box: mean |bias| per axis ~ scale-averaged voxel / 2
    assert 0.2 * scale < np.abs(gen / vox - (g["refine_helper"] / vox)[j]).mean() * 4 < 2.0 * scale


def test_chamfer_fixture_and_batch():
    g = _g()
    pred, gt = g["cd_pred"], g["cd_gt"]
    cd = postproc.chamfer_distance(torch.from_numpy(pred).to(DEV), torch.tensor([len(pred)], device=DEV),
                                   torch.from_numpy(gt).to(DEV))
    assert cd.shape == (1, 3) and cd.dtype == torch.float64
    ref = float(g["cd_value"])
    assert abs(float(cd[0, 0]) - ref) <= 1e-6 * ref
    assert abs(float(cd[0, 0]) - 0.5 * (float(cd[0, 1]) + float(cd[0, 2]))) < 1e-12
    # ragged batch: per-frame counts, padded rows must be ignored; one empty prediction -> inf
    rs = np.random.RandomState(3)
    B, cap_p, cap_g = 5, 5000, 2600
    P = rs.randn(B, cap_p, 3).astype(np.float32) * [4, 6, 1.5]
    G = rs.randn(B, cap_g, 3).astype(np.float32) * [4, 6, 1.5]
    pc = np.asarray([5000, 1, 0, 2049, 1024], np.int32)
    gc = np.asarray([2600, 2600, 10, 1, 2048], np.int32)
    out = postproc.chamfer_distance(torch.from_numpy(P.astype(np.float32)).to(DEV), torch.from_numpy(pc).to(DEV),
                                    torch.from_numpy(G.astype(np.float32)).to(DEV), torch.from_numpy(gc).to(DEV)).cpu().numpy()
    for b in range(B):
        want = orc.chamfer_distance(P[b, :pc[b]].astype(np.float32), G[b, :gc[b]].astype(np.float32))
        if np.isinf(want):
            assert np.isinf(out[b, 0])
        else:
            assert abs(out[b, 0] - want) <= 1e-6 * want, (b, out[b], want)
    # identical clouds: exactly zero
    same = postproc.chamfer_distance(torch.from_numpy(gt).to(DEV), torch.tensor([len(gt)], device=DEV),
                                     torch.from_numpy(gt).to(DEV))
    assert float(same[0, 0]) == 0.0


def test_chamfer_on_compacted_occupancy():
    """The metric consumes occupied_points' output directly (device counts, padded rows), as evaluate() chains them
    (engine_generation.py:283-319): threshold -> inverse norm -> polar2cartesian -> cal_metrics."""
    rs = np.random.RandomState(11)
    B, Q = 3, 40000
    q = (rs.rand(B, Q, 3).astype(np.float32) * 2 - 1)
    lg = rs.randn(B, Q).astype(np.float32) - 1.2
    gtn = (rs.rand(B, 3000, 3).astype(np.float32) * 2 - 1)
    rng = [0, -90, -20, 15.8, 90, 20]
    pts, cnt, _ = postproc.occupied_points(torch.from_numpy(lg).to(DEV), torch.from_numpy(q).to(DEV), pc_range=rng,
                                           view_cone=True)
    sc, off = orc.norm_constants(rng)
    gt_polar = torch.from_numpy(gtn).to(DEV) * torch.tensor(sc, device=DEV) + torch.tensor(off, device=DEV)
    gt_cart = postproc.polar_to_cartesian(gt_polar)
    cd = postproc.chamfer_distance(pts, cnt, gt_cart).cpu().numpy()
    for b in range(B):
        pred = orc.occupancy_points(lg[b], q[b], pc_range=rng, view_cone=True)
        gt = orc.occupancy_points(np.ones(3000, np.float32), gtn[b], pc_range=rng, view_cone=True)
        want = orc.chamfer_distance(pred, gt)
        assert abs(cd[b, 0] - want) <= 2e-5 * want, (cd[b], want)     # fp32 sin/cos of the two sides differ in the ulp


def test_radar_cube_prep_fixture_bit_exact():
    g = _g()
    ni, mi, nd, md, up, ta, te = g["radar_cfg"].tolist()
    cfg = dict(norm_intensity=bool(ni), max_intensity=mi, norm_dopp=bool(nd), max_dopp=md, upsample=bool(up),
               tgt_r_dim=128, tgt_a_dim=int(ta), tgt_e_dim=int(te))
    raw = torch.from_numpy(g["radar_raw"]).to(DEV)
    out = radar_prep.process_radar_data(raw, cfg)
    assert out.shape == (1, 128, 64, 32, 2)
    assert np.array_equal(out.cpu().numpy(), g["radar_processed"])
    # single un-batched cube, intensity channel only (the encoder's input), early return (no upsample, raw doppler)
    one = radar_prep.process_radar_data(raw[0], cfg, channels_out=1)
    assert np.array_equal(one.cpu().numpy()[..., 0], g["radar_processed"][0][..., 0])
    early = radar_prep.process_radar_data(raw[0], cfg, early_return=True)
    want = orc.process_radar_data(g["radar_raw"][0], bool(ni), mi, False, 1.0, False, 0, 0)
    assert np.array_equal(early.cpu().numpy(), want)
    # other target sizes, seeded, against the oracle
    rs = np.random.RandomState(5)
    raw2 = rs.uniform(-3, 50, (16, 5, 3, 4)).astype(np.float32)
    cfg2 = dict(cfg, tgt_r_dim=16, tgt_a_dim=17, tgt_e_dim=3)
    o2 = radar_prep.process_radar_data(torch.from_numpy(raw2).to(DEV), cfg2)
    assert np.array_equal(o2.cpu().numpy(), orc.process_radar_data(raw2, bool(ni), mi, bool(nd), md, True, 17, 3))


def test_prepared_cube_feeds_the_denoiser_conditioning():
    """raw cube -> process_radar_data (device) -> EDMPrecond.process_radar_cond equals the path through the
    reference-prepared cube of the fixture (the cubes themselves are bit-identical, see above; the encoder's GroupNorm
    statistics are accumulated with fp64 atomics across CTAs, so two runs may differ in the last bit)."""
    from helpers import build_denoiser, rel_l2
    g = _g()
    ni, mi, nd, md, up, ta, te = g["radar_cfg"].tolist()
    cfg = dict(norm_intensity=bool(ni), max_intensity=mi, norm_dopp=bool(nd), max_dopp=md, upsample=bool(up),
               tgt_r_dim=128, tgt_a_dim=int(ta), tgt_e_dim=int(te))
    net = build_denoiser(device=DEV)
    cube = radar_prep.process_radar_data(torch.from_numpy(g["radar_raw"]).to(DEV), cfg)
    a = net.process_radar_cond(cube)
    b = net.process_radar_cond(torch.from_numpy(g["radar_processed"]).to(DEV))
    assert a.shape == (1, 64, 512) and rel_l2(a, b) < 1e-3


@torch.no_grad()
def test_refine_pass_end_to_end():
    """evaluate()'s refine branch chained on the device: first pass -> refine_pass -> Chamfer, against the same chain
    built from the individually parity-tested pieces (and the oracle for the final compaction)."""
    from helpers import build_ae
    from rald_b200 import synth
    ae = build_ae(device=DEV)
    rng, vox = [0, -90, -20, 15.8, 90, 20], [0.05, 0.25, 0.5]
    z = torch.randn(2, 512, 32, generator=torch.Generator().manual_seed(3)).to(DEV)
    q = synth.query_points(2, 8192, seed=17).to(DEV)
    lg = ae.decode(z, q).squeeze(-1)
    thr = float(torch.quantile(lg, 0.9, dim=1).min())   # random-init logits: the frames' common modes differ
    pts, cnt, _ = postproc.occupied_points(lg, q, threshold=thr, pc_range=rng)
    assert int(cnt.min()) > 0
    pts2, cnt2, rq = postproc.refine_pass(ae, z, pts, cnt, 20000, rng, vox, aug_scale=10, threshold=thr, view_cone=True,
                                          rng="device", seed=9)
    assert rq.shape == (2, 20000, 3) and float(rq.abs().max()) <= 1.0
    lg2 = ae.decode(z, rq).squeeze(-1)
    for b in range(2):
        want = orc.occupancy_points(lg2[b].cpu().numpy(), rq[b].cpu().numpy(), pc_range=rng, view_cone=True,
                                    threshold=thr)
        assert int(cnt2[b]) == len(want)
        assert np.allclose(pts2[b, :len(want)].cpu().numpy(), want, rtol=0, atol=2e-5)
    cd = postproc.chamfer_distance(pts2, cnt2, postproc.polar_to_cartesian(pts[:, :4000]), cnt.clamp(max=4000))
    assert torch.isfinite(cd).all() and float(cd.min()) >= 0.0


@torch.no_grad()
def test_generate_point_clouds_pipeline():
    """evaluate()'s per-batch body chained on the device equals the same chain assembled by hand from the
    parity-tested pieces; the host query grid replays the reference's np.random.uniform draws."""
    from helpers import build_ae, build_denoiser
    from rald_b200 import evaluate, synth
    rng = [0, -90, -20, 15.8, 90, 20]
    np.random.seed(5)
    g = evaluate.generate_query_points(1000, rng)
    np.random.seed(5)
    want = np.stack([np.random.uniform(-1, 1, 1000) for _ in range(3)], axis=1)
    assert g.dtype == np.float64 and np.array_equal(g, want)
    net, ae = build_denoiser(device=DEV), build_ae(device=DEV)
    cube = synth.radar_cube(2, seed=21).to(DEV)
    grid = synth.query_points(1, 16384, seed=4)[0].to(DEV)
    z = net.sample(cond=cube, batch_seeds=None, cond_type="radar")
    lg = ae.decode(z, grid[None].expand(2, -1, -1).contiguous()).squeeze(-1)
    thr = float(torch.quantile(lg, 0.9, dim=1).min())
    gt = synth.query_points(2, 3000, seed=8).to(DEV)
    out = evaluate.generate_point_clouds(net, ae, cube, 16384, rng, threshold=thr, ground_truth=gt, grid=grid)
    assert torch.equal(out["latents"], z)
    pts, cnt, _ = postproc.occupied_points(lg, grid[None].expand(2, -1, -1).contiguous(), thr, rng, view_cone=True)
    assert torch.equal(out["counts"], cnt) and int(cnt.min()) > 0
    n = int(cnt.max())
    assert torch.equal(out["points"][:, :n], pts[:, :n]) or all(
        torch.equal(out["points"][b, :int(cnt[b])], pts[b, :int(cnt[b])]) for b in range(2))
    for b in range(2):
        pred = orc.occupancy_points(lg[b].cpu().numpy(), grid.cpu().numpy(), pc_range=rng, view_cone=True, threshold=thr)
        ref_gt = orc.occupancy_points(np.ones(3000, np.float32), gt[b].cpu().numpy(), pc_range=rng, view_cone=True)
        want_cd = orc.chamfer_distance(pred, ref_gt)
        assert abs(float(out["cd"][b]) - want_cd) <= 2e-5 * want_cd
    # with the refine pass: shapes / ranges / finite metric
    out2 = evaluate.generate_point_clouds(net, ae, cube, 16384, rng, threshold=thr, ground_truth=gt, grid=grid,
                                          refine=dict(aug_num=30000, voxel_size=[0.05, 0.25, 0.5], scale=10, seed=3))
    assert out2["points"].shape[0] == 2 and torch.isfinite(out2["cd"]).all()
