"""Adds tests/golden/churn.npz: the UNMODIFIED reference edm_sampler with S_churn > 0 (the stochastic branch,
models_radar_generation.py:254-260) on the default denoiser, 4 steps, the per-step noise drawn by a seeded generator
and recorded so that the GPU test can inject the same draws.

    python tests/golden/make_golden_churn.py        # ~25 s, needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from make_golden import SEED, build_denoiser  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402

NUM_STEPS, S_CHURN, S_NOISE = 4, 2.0, 1.003


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    _, m_gen, _ = ref_import.import_reference()
    cfg = ref_import.load_generation_config()
    net = build_denoiser(m_gen.__dict__, cfg.ar_model.name, cfg.ar_model.configs)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    cube = synth.radar_cube(2, seed=SEED)
    lat = synth.unit_latents([0, 1])
    gen = torch.Generator("cpu").manual_seed(77)
    noises = []

    def randn_like(x):
        n = torch.randn(x.shape, generator=gen, dtype=x.dtype)
        noises.append(n)
        return n

    x_ref = m_gen.edm_sampler(net, lat, cube, "radar", randn_like=randn_like, num_steps=NUM_STEPS, S_churn=S_CHURN,
                              S_noise=S_NOISE)
    assert len(noises) == NUM_STEPS
    tok = net.process_radar_cond(cube)
    x_orc = orc.edm_sample(sd, lat, tok, num_steps=NUM_STEPS, S_churn=S_CHURN, S_noise=S_NOISE, noises=noises)
    e = orc.rel_l2(x_orc, x_ref)
    print("oracle vs reference (S_churn > 0) rel-L2", e)
    assert e < 1e-4
    np.savez_compressed(os.path.join(HERE, "churn.npz"), x=x_ref.numpy(), noises=torch.stack(noises).numpy(),
                        num_steps=np.int32(NUM_STEPS), s_churn=np.float32(S_CHURN),
                        s_noise=np.float32(S_NOISE))


if __name__ == "__main__":
    main()
