"""Fixture tests/golden/trainloop.npz for SURVEY.md §8(f) rows 3 / 4 helpers, produced by UNMODIFIED reference code:

  * update_ema     engine_generation.py:29-39 — the module imports open3d / spconv-dependent datasets, absent here,
                   so the FUNCTION's source is cut out of the file with `ast` and compiled unchanged against torch.
  * accuracy / IoU engine_generation.py:376-385 — statements inside cache_latents' loop body; the statement range
                   is located with `ast` (from `threshold = 0` to the `iou = iou.mean()` line) and executed unchanged on
                   synthetic `outputs` / `labels`, except that the two batch-mean statements are skipped so that the
                   per-frame values are kept.
Asserts that the oracle restatements reproduce them bit for bit, then commits inputs + outputs.

    python tests/golden/make_golden_trainloop.py        # a second, needs /root/reference
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402


def main():
    path = os.path.join(ref_import.REF_ROOT, "engine_generation.py")
    src = open(path).read()
    tree = ast.parse(src)
    out = {}

    # ---- update_ema ----
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "update_ema")
    ns = {"torch": torch}
    exec(compile(ast.get_source_segment(src, fn), "engine_generation.update_ema", "exec"), ns)
    g = torch.Generator().manual_seed(5)
    shapes = [(4097,), (3, 5), (1,), (2, 4096), (8192,), (33, 7, 3)]
    rate = 0.9999                                            # main_generation.py --ema_rate style value
    targ = [torch.randn(s, generator=g) for s in shapes]
    srcs = [torch.randn(s, generator=g) for s in shapes]
    ref_t = [t.clone() for t in targ]
    for _ in range(3):
        ns["update_ema"](ref_t, srcs, rate=rate)
    orc_t = [t.clone().numpy() for t in targ]
    for _ in range(3):
        orc.update_ema(orc_t, [s.numpy() for s in srcs], rate=rate)
    for a, b in zip(ref_t, orc_t):
        assert np.array_equal(a.numpy(), b), "update_ema oracle != reference"
    ref_default = [t.clone() for t in targ]
    ns["update_ema"](ref_default, srcs)                      # default rate 0.99
    for i, s in enumerate(shapes):
        out[f"ema_target_{i}"] = targ[i].numpy()
        out[f"ema_source_{i}"] = srcs[i].numpy()
        out[f"ema_after3_{i}"] = ref_t[i].numpy()
        out[f"ema_default_{i}"] = ref_default[i].numpy()
    out["ema_rate"] = np.float64(rate)

    # ---- accuracy / IoU statements of cache_latents ----
    cl = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "cache_latents")
    loop = next(n for n in cl.body if isinstance(n, ast.For))
    stmts = []
    on = False
    for st in loop.body:
        seg = ast.get_source_segment(src, st)
        if seg.startswith("threshold ="):
            on = True
        if on and not seg.endswith(".mean()"):
            stmts.append(seg)
        if seg.startswith("iou = iou.mean()"):
            break
    assert len(stmts) == 7, stmts
    B, Q = 5, 10007
    outputs = torch.randn(B, Q, generator=g) * 0.5 - 0.2
    outputs[0, :17] = 0.0                                    # logits exactly at the threshold count as occupied
    labels = (torch.rand(B, Q, generator=g) < 0.3).float()
    labels[4] = 0.0
    outputs[4] = -1.0                                        # empty union -> NaN + 1e-5
    env = {"torch": torch, "outputs": outputs, "labels": labels}
    exec("\n".join(stmts), env)
    acc, iou = orc.occupancy_iou(outputs, labels, 0.0)
    assert np.array_equal(env["accuracy"].numpy(), acc.numpy())
    assert np.array_equal(env["iou"].numpy(), iou.numpy(), equal_nan=True)
    out["iou_logits"] = outputs.numpy()
    out["iou_labels"] = labels.numpy().astype(np.uint8)
    out["iou_accuracy"] = env["accuracy"].numpy()
    out["iou_iou"] = env["iou"].numpy()

    np.savez_compressed(os.path.join(HERE, "trainloop.npz"), **out)
    print("wrote trainloop.npz:", {k: v.shape for k, v in out.items() if k.startswith("iou")})


if __name__ == "__main__":
    main()
