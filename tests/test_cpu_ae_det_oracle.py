"""CPU: the deterministic AutoEncoder (reference models_ae.py:181-282, factory ae_d512_m512) — the package's module has
the reference's seeded state_dict (hash from the unmodified reference in tests/golden/ae_det.npz) and the oracle
restatement reproduces the reference's latents and logits on the fixture inputs."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import cpu_state_dict, sd_hash
from oracle import rald_oracle as orc
from rald_b200 import _lib, models_ae


def _g():
    return np.load(os.path.join(GOLDEN, "ae_det.npz"))


def _build(n):
    torch.manual_seed(1024)
    return models_ae.ae_d512_m512(N=n).eval()


def test_state_dict_is_the_references():
    g = _g()
    ae = _build(int(g["pc"].shape[1]))
    sd = cpu_state_dict(ae)
    assert len(sd) == int(g["n_keys"]) and sd_hash(sd) == str(g["state_hash"])
    assert not any(k.startswith(("proj.", "mean_fc.", "logvar_fc.")) for k in sd)


def test_oracle_matches_reference_fixture():
    g = _g()
    ae = _build(int(g["pc"].shape[1]))
    sd = cpu_state_dict(ae)
    pc, q = torch.from_numpy(g["pc"]), torch.from_numpy(g["queries"])
    assert np.array_equal(orc.fps_indices(pc, 512)[0].numpy(), g["fps_idx"])
    x = orc.ae_encode_stats(sd, pc, "point", 512)
    assert np.allclose(x[0, ::8].numpy(), g["latents_rows"], rtol=0, atol=1e-6)
    logits = orc.ae_decode(sd, x, q)[0, :, 0]
    assert np.allclose(logits.numpy(), g["logits"], rtol=0, atol=1e-6)


def test_no_cpu_path_and_unsupported_geometry():
    ae = _build(2048)
    with pytest.raises(_lib.RaldError):
        ae(torch.zeros(1, 2048, 3), torch.zeros(1, 8, 3))
    assert isinstance(models_ae.ae_d512_m256(N=2048), models_ae.AutoEncoder)   # constructs; kernels need 512 latents
