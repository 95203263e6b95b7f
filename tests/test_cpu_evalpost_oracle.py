"""CPU: the oracle restatements of the eval-loop helpers against the fixture written from
the UNMODIFIED reference functions (tests/golden/make_golden_evalpost.py): aug_query_helper + norm_points,
cal_metrics, Coloradar_dataset.process_radar_data. Bit-exact except the Chamfer value (fp64 sums, 1e-12)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import rald_oracle as orc
from rald_b200 import postproc


def _g():
    return np.load(os.path.join(GOLDEN, "evalpost.npz"))


def test_refine_queries_oracle_bit_exact():
    g = _g()
    aug = orc.aug_query_helper(g["refine_helper"], int(g["refine_aug_num"]), g["pc_range"].tolist(),
                               g["voxel_size"].tolist(), int(g["refine_scale"]), g["refine_sel"], g["refine_scales"],
                               g["refine_u"])
    q = orc.norm_points(aug, g["pc_range"].tolist())
    assert q.dtype == np.float32 and np.array_equal(q, g["refine_queries"])
    # every generated point stays inside the range, i.e. inside the normalised cube
    assert np.abs(q).max() <= 1.0


def test_refine_draw_order_is_the_references():
    """draw_refine_randoms issues the reference's np.random calls in its order: replaying the fixture's seed gives the
    fixture's draws."""
    g = _g()
    np.random.seed(1234)
    sel, scales, u = postproc.draw_refine_randoms(g["refine_helper"].shape[0], int(g["refine_aug_num"]),
                                                  int(g["refine_scale"]))
    assert np.array_equal(sel, g["refine_sel"]) and np.array_equal(scales, g["refine_scales"])
    assert np.array_equal(u, g["refine_u"])
    assert postproc.draw_refine_randoms(10, 10, 3) is None


def test_chamfer_oracle():
    g = _g()
    assert abs(orc.chamfer_distance(g["cd_pred"], g["cd_gt"]) - float(g["cd_value"])) < 1e-12
    assert orc.chamfer_distance(g["cd_pred"][:0], g["cd_gt"]) == float("inf")


def test_radar_cube_prep_oracle_bit_exact():
    g = _g()
    ni, mi, nd, md, up, ta, te = g["radar_cfg"].tolist()
    for raw, ref in zip(g["radar_raw"], g["radar_processed"]):
        o = orc.process_radar_data(raw, bool(ni), mi, bool(nd), md, bool(up), int(ta), int(te))
        assert o.dtype == np.float32 and np.array_equal(o, ref)
        # grid nodes of the align_corners=True upsample reproduce the normalised input exactly
        assert np.array_equal(o[:, ::9, ::31, 0], np.clip(raw[..., 0], 0, mi) / np.float32(mi))
