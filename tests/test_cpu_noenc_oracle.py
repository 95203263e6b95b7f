"""CPU: the encoder-less conditioning variant (`use_radar_enc: false`: 2048 raw-cube tokens) — oracle and seeded
drop-in initialisation against the fixture written from the unmodified reference (tests/golden/make_golden_noenc.py)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from helpers import SEED, cpu_state_dict, rel_l2, sd_hash
from oracle import rald_oracle as orc
from rald_b200 import models_radar_generation, synth
from rald_b200.config import DEFAULT_DENOISER_NAME, default_denoiser_configs


def build_noenc(device="cpu"):
    cfg = default_denoiser_configs()
    cfg.use_radar_enc = False
    cfg.unfreeze_radar_enc = False
    torch.manual_seed(SEED)
    net = models_radar_generation.__dict__[DEFAULT_DENOISER_NAME](configs=cfg).eval()
    net.model.proj_out.reset_parameters()
    return net.to(device)


@torch.no_grad()
def test_noenc_oracle_and_seeded_init():
    g = np.load(os.path.join(GOLDEN, "noenc.npz"))
    net = build_noenc()
    sd = cpu_state_dict(net)
    assert sd_hash(sd) == bytes(g["hash"]).decode()
    assert "radar_enc.conv_in.weight" not in sd and sd["radar_token_project.weight"].shape == (512, 1)
    cube = torch.from_numpy(g["cube"])
    tokens = orc.process_radar_cond(sd, cube, use_encoder=False)
    assert tokens.shape == (1, 2048, 512)
    assert rel_l2(tokens[:, :64], torch.from_numpy(g["tokens_head"])) < 1e-6
    s = torch.tensor(1.5)
    d = orc.edm_precond(sd, synth.unit_latents([0]) * s, s, tokens)
    assert rel_l2(d, torch.from_numpy(g["denoised_1.5"])) < 2e-5
