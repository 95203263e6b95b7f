"""CPU check that pins the oracle's BACKWARD to the reference: autograd through oracle/rald_oracle.py (the functional
restatement) on the fixture's inputs reproduces the loss, D and parameter gradients the UNMODIFIED reference produced
(tests/golden/train_grads.npz, written by make_golden_train.py from /root/reference)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from helpers import build_denoiser, cpu_state_dict, grad_sample_index, rel_l2
from oracle import rald_oracle as orc
from rald_b200 import synth


def test_oracle_autograd_reproduces_reference_gradients():
    fx = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    sd = cpu_state_dict(build_denoiser())
    cube = synth.radar_cube(2, seed=1024)
    y, sigma, noise = (torch.from_numpy(fx[k]) for k in ("y", "sigma", "noise"))
    with torch.no_grad():
        feat_sd = {k: v for k, v in sd.items()}
        tok_const = orc.process_radar_cond(feat_sd, cube)
    for k, v in sd.items():
        if not k.startswith("radar_enc."):
            v.requires_grad_(True)
    # the token projection / embeddings stay differentiable: redo the (cheap) tail of process_radar_cond under autograd
    tok = orc.process_radar_cond(sd, cube)
    assert rel_l2(tok, tok_const) < 1e-6
    D = orc.edm_precond(sd, y + noise * sigma, sigma, tok)
    weight = (sigma ** 2 + 1.0) / sigma ** 2
    loss = (weight * (D - y) ** 2).mean()
    loss.backward()
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    assert rel_l2(D, torch.from_numpy(fx["D"])) <= 1e-5
    names = [str(n) for n in fx["names"]]
    assert len(names) == 493
    for name in names:
        gr = sd[name].grad
        assert gr is not None, name
        assert abs(float(gr.double().norm()) - float(fx["norm/" + name])) <= 2e-4 * float(fx["norm/" + name]), name
        if "full/" + name in fx.files:
            assert rel_l2(gr, torch.from_numpy(fx["full/" + name])) <= 2e-4, name
        else:
            idx = torch.from_numpy(grad_sample_index(name, gr.numel()))
            assert rel_l2(gr.reshape(-1)[idx], torch.from_numpy(fx["sample/" + name])) <= 2e-4, name


def test_oracle_autograd_reproduces_reference_encoder_gradients():
    """The same step with the radar encoder trainable (the shipped configuration): autograd through the oracle's encoder
    restatement reproduces the 144 encoder gradients of tests/golden/train_grads_enc.npz."""
    fx0 = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    fx = np.load(os.path.join(GOLDEN, "train_grads_enc.npz"))
    sd = cpu_state_dict(build_denoiser())
    for v in sd.values():
        v.requires_grad_(True)
    cube = synth.radar_cube(2, seed=1024)
    y, sigma, noise = (torch.from_numpy(fx0[k]) for k in ("y", "sigma", "noise"))
    tok = orc.process_radar_cond(sd, cube)
    D = orc.edm_precond(sd, y + noise * sigma, sigma, tok)
    loss = ((sigma ** 2 + 1.0) / sigma ** 2 * (D - y) ** 2).mean()
    loss.backward()
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    names = [str(n) for n in fx["names"]]
    assert len(names) == 144
    for name in names:
        gr = sd[name].grad
        assert gr is not None, name
        if name.endswith(".k.bias"):       # analytically zero: rounding noise on both sides
            continue
        assert abs(float(gr.double().norm()) - float(fx["norm/" + name])) <= 1e-3 * float(fx["norm/" + name]), name
        if "full/" + name in fx.files:
            assert rel_l2(gr, torch.from_numpy(fx["full/" + name])) <= 1e-3, name
        else:
            idx = torch.from_numpy(grad_sample_index(name, gr.numel()))
            assert rel_l2(gr.reshape(-1)[idx], torch.from_numpy(fx["sample/" + name])) <= 1e-3, name
