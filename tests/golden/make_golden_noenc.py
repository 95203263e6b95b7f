"""Fixture tests/golden/noenc.npz: the UNMODIFIED reference EDMPrecond built WITHOUT the radar encoder
(`use_radar_enc: false`, `unfreeze_radar_enc: false`: the raw [128, 8, 2] intensity cube becomes 2048 conditioning
tokens, models_radar_generation.py:357-361, 378-405) evaluated once per sigma on a seeded input; asserts the oracle
reproduces it. The long-context (chunked) cross-attention path of rald_b200 is checked against it.

    python tests/golden/make_golden_noenc.py        # ~1 min, needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from make_golden import SEED, build_denoiser, sd_hash  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402
from rald_b200.config import DEFAULT_DENOISER_NAME  # noqa: E402


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    _, m_gen, _ = ref_import.import_reference()
    cfg = ref_import.load_generation_config().ar_model.configs
    cfg = ref_import.EasyDict(dict(cfg))
    cfg.use_radar_enc = False
    cfg.unfreeze_radar_enc = False
    net = build_denoiser(m_gen.__dict__, DEFAULT_DENOISER_NAME, cfg)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(77)
    cube = torch.rand(1, cfg.input_radar_r_dim, cfg.input_radar_a_dim, cfg.input_radar_e_dim, 2, generator=g)
    lat = synth.unit_latents([0])
    out = {"cube": cube.numpy(), "hash": np.frombuffer(sd_hash(sd).encode(), dtype=np.uint8)}
    tokens = net.process_radar_cond(cube)
    assert tokens.shape == (1, 2048, 512)
    e = orc.rel_l2(orc.process_radar_cond(sd, cube, use_encoder=False), tokens)
    print("oracle vs reference tokens rel-L2", e)
    assert e < 1e-6
    out["tokens_head"] = tokens[:, :64].numpy()     # the first 64 of the 2048 tokens (the oracle recomputes the rest)
    for sigma in (80.0, 1.5):
        s = torch.tensor(sigma)
        d = net(lat * s, s, cube, cond_type="radar")
        e = orc.rel_l2(orc.edm_precond(sd, lat * s, s, tokens), d)
        print(f"oracle vs reference denoised sigma={sigma}: rel-L2 {e:.3e}")
        assert e < 2e-5
        out[f"denoised_{sigma}"] = d.numpy()
    path = os.path.join(HERE, "noenc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
