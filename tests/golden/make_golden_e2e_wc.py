"""Adds the WELL-CONDITIONED end-to-end fixture tests/golden/e2e_wc.npz.

With constructor-random weights the occupancy field of the reference is almost constant (std 1.5e-3 on a -0.165
offset, nothing above 0), so `logit > 0` (engine_generation.py:285) measures nothing but a common-mode offset. Here the
SAME seeded state_dict is sharpened on BOTH sides (the transform is part of the shared state_dict, not of either
implementation):

    decoder_cross_attn.fn.to_q.weight *= 32     (softmax over the 512 latents no longer near-uniform)
    to_outputs.weight                 *= 8      (field std ~0.3: logits O(0.1 - 1), as with trained weights)
    to_outputs.bias                    = value that puts the reference's 95-th percentile logit at 0

and the UNMODIFIED reference KLAutoEncoder decodes the reference sampler's final latents (sampler_trace.npz) at 32768
uniform queries. The test then applies the literal `logit > 0` rule to both sides.

    python tests/golden/make_golden_e2e_wc.py        # ~15 s, needs /root/reference
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

Q_SCALE, OUT_SCALE = 32.0, 8.0


@torch.no_grad()
def main():
    m_ae, _, _ = ref_import.import_reference()
    torch.manual_seed(1024)
    ae = m_ae.kl_d512_m512_l32_mix(N=10000).eval()
    ae.decoder_cross_attn.fn.to_q.weight.mul_(Q_SCALE)
    ae.to_outputs.weight.mul_(OUT_SCALE)
    z = torch.from_numpy(np.load(os.path.join(HERE, "sampler_trace.npz"))["trace"][-1])[None]   # [1, 512, 32]
    q = synth.query_points(1, 32768, seed=99)
    raw = ae.decode(z, q)[..., 0]
    p95 = float(np.quantile(raw.numpy(), 0.95))
    ae.to_outputs.bias.sub_(p95)
    bias = float(ae.to_outputs.bias[0])
    logits = ae.decode(z, q)[..., 0]
    sd = {k: v.detach() for k, v in ae.state_dict().items()}
    e = orc.rel_l2(orc.ae_decode(sd, z, q)[..., 0], logits)
    print("oracle vs reference logits rel-L2", e)
    assert e < 1e-4
    occ = float((logits > 0).float().mean())
    print("logit mean %.4f std %.4f min %.3f max %.3f; occupied fraction %.4f; bias %.6f"
          % (float(logits.mean()), float(logits.std()), float(logits.min()), float(logits.max()), occ, bias))
    assert 0.03 < occ < 0.07 and float(logits.std()) > 0.1
    np.savez_compressed(os.path.join(HERE, "e2e_wc.npz"), logits=logits.numpy(), q_scale=np.float32(Q_SCALE),
                        out_scale=np.float32(OUT_SCALE), bias=np.float32(bias))


if __name__ == "__main__":
    main()
