"""Adds the end-to-end fixture tests/golden/e2e.npz: the UNMODIFIED reference KLAutoEncoder (kl_d512_m512_l32_mix,
seed 1024) decodes the reference sampler's final latents (sampler_trace.npz, written by make_golden.py from the
reference's own edm_sampler) at 32768 uniform query points. Used by the Chamfer criterion of the north star:
point clouds = queries with logit > 0 after the harness convention of SURVEY.md §7.3 (the reference's 95-th
percentile logit is subtracted from to_outputs.bias on BOTH sides, because random-init logits are all negative).

    python tests/golden/make_golden_e2e.py        # ~10 s, needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402


@torch.no_grad()
def main():
    m_ae, _, _ = ref_import.import_reference()
    torch.manual_seed(1024)
    ae = m_ae.kl_d512_m512_l32_mix(N=10000).eval()
    z = torch.from_numpy(np.load(os.path.join(HERE, "sampler_trace.npz"))["trace"][-1])[None]   # [1, 512, 32]
    q = synth.query_points(1, 32768, seed=99)
    logits = ae.decode(z, q)[..., 0]
    sd = {k: v.detach() for k, v in ae.state_dict().items()}
    e = orc.rel_l2(orc.ae_decode(sd, z, q)[..., 0], logits)
    print("oracle vs reference e2e logits rel-L2", e)
    assert e < 1e-4
    shift = float(np.quantile(logits.numpy(), 0.95))
    np.savez_compressed(os.path.join(HERE, "e2e.npz"), logits=logits.numpy(), shift=np.float32(shift))
    print("logit mean %.5f std %.5f, 95-th percentile %.5f" % (float(logits.mean()), float(logits.std()), shift))


if __name__ == "__main__":
    main()
