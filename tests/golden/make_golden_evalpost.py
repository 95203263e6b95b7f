"""Fixture tests/golden/evalpost.npz for SURVEY.md §8(f) rows 1, 2, 4, produced by the UNMODIFIED reference functions:

  * aug_query_helper            datasets/utils/query_helper.py   (imported by file path; numpy only)
  * norm_points, inverse_norm_points, cal_metrics   utils/utils.py   (imported by file path; numpy / scipy / torch)
  * Coloradar_dataset.process_radar_data   datasets/aligned_coloradar/Coloradar_dataset.py — the module imports
    spconv / open3d, which are absent here, so the METHOD's source is cut out of the file with `ast` and compiled
    unchanged against numpy / torch / F; `self` is a stand-in that only carries `.config`.

Asserts that the oracle restatements (oracle/rald_oracle.py) reproduce them, then commits inputs + outputs.

    python tests/golden/make_golden_evalpost.py        # a few seconds, needs /root/reference
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _method_source(path, func):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            seg = ast.get_source_segment(src, node)
            lines = src.splitlines()[node.lineno - 1:node.end_lineno]
            indent = len(lines[0]) - len(lines[0].lstrip())
            return "\n".join(l[indent:] for l in lines), seg
    raise KeyError(func)


def main():
    root = ref_import.REF_ROOT
    cfg = ref_import.load_generation_config()
    qh = _load(os.path.join(root, "datasets/utils/query_helper.py"), "ref_query_helper")
    ut = _load(os.path.join(root, "utils/utils.py"), "ref_utils")
    text, _ = _method_source(os.path.join(root, "datasets/aligned_coloradar/Coloradar_dataset.py"), "process_radar_data")
    ns = {"np": np, "torch": torch, "F": F}
    exec(compile(text, "Coloradar_dataset.process_radar_data", "exec"), ns)
    process_radar_data = ns["process_radar_data"]

    lidar, radar = cfg.dataset.lidar, cfg.dataset.radar
    inf = cfg.eval.inference
    out = {}

    # ---- refine_query: aug_query_helper + norm_points (engine_generation.py:291-297) ----
    rs = np.random.RandomState(7)
    lo, hi = np.asarray(lidar.pc_range[:3], np.float32), np.asarray(lidar.pc_range[3:], np.float32)
    helper = (lo + (hi - lo) * rs.rand(1500, 3)).astype(np.float32)   # "pred": inverse-normalised polar points
    aug_num = 6000
    np.random.seed(1234)
    ref_aug = qh.aug_query_helper(helper.copy(), aug_num, lidar.pc_range, lidar.voxel_size, inf.refine_query_scale)
    ref_q = ut.norm_points(ref_aug, lidar.pc_range, lidar.norm_anisotropy, lidar.norm_isotropy)
    # the same draws, replayed in the reference's order
    np.random.seed(1234)
    G = aug_num - helper.shape[0]
    sel = np.random.choice(helper.shape[0], size=G, replace=True)
    scales = np.random.choice(np.arange(inf.refine_query_scale, step=1) + 1, size=G)
    u = np.random.rand(G, 3)
    o_aug = orc.aug_query_helper(helper, aug_num, lidar.pc_range, lidar.voxel_size, inf.refine_query_scale, sel, scales, u)
    o_q = orc.norm_points(o_aug, lidar.pc_range, lidar.norm_anisotropy, lidar.norm_isotropy)
    assert ref_aug.dtype == np.float32 and np.array_equal(ref_aug, o_aug), "aug_query_helper oracle != reference"
    assert ref_q.dtype == np.float32 and np.array_equal(ref_q, o_q), "norm_points oracle != reference"
    # truncation branch (N >= aug_num)
    ref_trunc = qh.aug_query_helper(helper.copy(), 1000, lidar.pc_range, lidar.voxel_size, inf.refine_query_scale)
    assert np.array_equal(ref_trunc, orc.aug_query_helper(helper, 1000, lidar.pc_range, lidar.voxel_size,
                                                          inf.refine_query_scale, None, None, None))
    out.update(refine_helper=helper, refine_sel=sel.astype(np.int32), refine_scales=scales.astype(np.int32),
               refine_u=u, refine_queries=ref_q, refine_aug_num=np.int64(aug_num),
               refine_scale=np.int64(inf.refine_query_scale), pc_range=np.asarray(lidar.pc_range, np.float64),
               voxel_size=np.asarray(lidar.voxel_size, np.float64))

    # ---- Chamfer: cal_metrics on two clouds (cartesian, metres) ----
    pred = (rs.rand(3000, 3) * [15.0, 20.0, 6.0] - [0.0, 10.0, 3.0]).astype(np.float32)
    gt = (rs.rand(2000, 3) * [15.0, 20.0, 6.0] - [0.0, 10.0, 3.0]).astype(np.float32)
    pred[:50] = gt[:50]                                   # exact coincidences (distance 0)
    cd = ut.cal_metrics(y_pred=pred, y_gt=gt)
    assert abs(cd - orc.chamfer_distance(pred, gt)) < 1e-12
    assert ut.cal_metrics(y_pred=pred[:0], y_gt=gt) == np.inf
    out.update(cd_pred=pred, cd_gt=gt, cd_value=np.float64(cd))

    # ---- radar cube prep: process_radar_data on raw cubes [128, 8, 2, 3] ----
    stand_in = types.SimpleNamespace(config=types.SimpleNamespace(radar=radar))
    raws, refs = [], []
    for k in range(1):
        raw = np.zeros((radar.input_r_dim, radar.input_a_dim, radar.input_e_dim, 3), np.float32)
        raw[..., 0] = rs.uniform(-5.0, 60.0, raw.shape[:3])            # dB, beyond the clip on both sides
        raw[..., 1] = rs.uniform(-2.5, 2.5, raw.shape[:3])
        raw[..., 2] = (rs.rand(*raw.shape[:3]) > 0.3).astype(np.float32)
        ref = process_radar_data(stand_in, raw.copy())
        assert ref.dtype == np.float32 and ref.shape == (radar.tgt_r_dim, radar.tgt_a_dim, radar.tgt_e_dim, 2)
        o = orc.process_radar_data(raw, radar.norm_intensity, radar.max_intensity, radar.norm_dopp, radar.max_dopp,
                                   radar.upsample, radar.tgt_a_dim, radar.tgt_e_dim)
        err = float(np.abs(o - ref).max())
        print("process_radar_data oracle vs reference: max abs %.3e, bit-exact %s" % (err, np.array_equal(o, ref)))
        assert np.array_equal(o, ref), err
        early = process_radar_data(stand_in, raw.copy(), early_return=True)
        assert np.array_equal(early, orc.process_radar_data(raw, radar.norm_intensity, radar.max_intensity, False, 1.0,
                                                            False, 0, 0))
        raws.append(raw)
        refs.append(ref)
    out.update(radar_raw=np.stack(raws), radar_processed=np.stack(refs),
               radar_cfg=np.asarray([float(radar.norm_intensity), radar.max_intensity, float(radar.norm_dopp),
                                     radar.max_dopp, float(radar.upsample), radar.tgt_a_dim, radar.tgt_e_dim], np.float64))
    path = os.path.join(HERE, "evalpost.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
