"""Imports the UNMODIFIED reference modules from /root/reference (this container only) for fixture generation.

Three third-party imports of the reference are absent from this image and are stubbed before import
(SURVEY.md §8c): timm's DropPath (identity in eval, parameter-free), torch_cluster.fps (restated in
oracle/rald_oracle.py:fps_indices, start index 0) and easydict.EasyDict. Nothing here is used at test time on
the GPU box: the fixtures it produces are committed under tests/golden/.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import yaml

REF_ROOT = os.environ.get("RALD_REFERENCE_ROOT", "/root/reference")
REPO_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
GEN_CFG = "configs/generation/ge_indoor_cfg_aniso_mix_view_cone_unfreeze_enc_ints_only_eval.yml"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


class EasyDict(dict):
    """Minimal stand-in: attribute access + recursive wrapping + .get()."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


def _install_stubs():
    if REPO_ROOT not in sys.path:
        sys.path.insert(0, REPO_ROOT)
    from oracle import rald_oracle

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training or self.drop_prob == 0.0, "stub DropPath is eval-only"
            return x

    timm = types.ModuleType("timm")
    timm.models = types.ModuleType("timm.models")
    timm.models.layers = types.ModuleType("timm.models.layers")
    timm.models.layers.DropPath = DropPath
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", timm.models)
    sys.modules.setdefault("timm.models.layers", timm.models.layers)

    def fps(src, batch, ratio=None, random_start=False):
        nb = int(batch.max().item()) + 1
        n = src.shape[0] // nb
        m = int(-(-ratio * n // 1))  # ceil(ratio * n)
        idx = rald_oracle.fps_indices(src.view(nb, n, 3), m)
        return (idx + (torch.arange(nb) * n)[:, None]).reshape(-1)

    tc = types.ModuleType("torch_cluster")
    tc.fps = fps
    sys.modules.setdefault("torch_cluster", tc)

    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules.setdefault("easydict", ed)


def import_reference():
    """Returns (models_ae, models_radar_generation, models_radar_encoder) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    m_ae = importlib.import_module("model.models_ae")
    m_gen = importlib.import_module("model.models_radar_generation")
    m_enc = importlib.import_module("model.models_radar_encoder")
    return m_ae, m_gen, m_enc


def load_generation_config():
    """The shipped eval YAML (configs/generation/...eval.yml) as an EasyDict."""
    with open(os.path.join(REF_ROOT, GEN_CFG)) as f:
        return EasyDict(yaml.safe_load(f))
