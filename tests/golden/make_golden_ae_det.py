"""Fixture tests/golden/ae_det.npz: the reference's deterministic ``AutoEncoder`` (models_ae.py:181-282, factory
``ae_d512_m512``) run UNMODIFIED on the CPU (import stubs of ref_import.py; torch_cluster.fps restated with start index
0) on a seeded cloud, plus the oracle restatement checked against it.

    python tests/golden/make_golden_ae_det.py        # ~1 min, needs /root/reference
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
import ref_import  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402

SEED, N, Q = 1024, 4096, 2048


def sd_hash(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    m_ae, _, _ = ref_import.import_reference()
    torch.manual_seed(SEED)
    ref = m_ae.__dict__["ae_d512_m512"](N=N).eval()
    sd = {k: v.detach().float() for k, v in ref.state_dict().items()}
    g = torch.Generator().manual_seed(77)
    pc = torch.rand(1, N, 3, generator=g) * 2 - 1
    q = torch.rand(1, Q, 3, generator=g) * 2 - 1
    with torch.no_grad():
        x = ref.encode(pc)
        logits = ref.decode(x, q)
        out = ref(pc, q)["logits"]
    assert torch.equal(out, logits.squeeze(-1))
    ox = orc.ae_encode_stats(sd, pc, "point", 512)
    ol = orc.ae_decode(sd, x, q)
    dx = float((ox - x).norm() / x.norm())
    dl = float((ol - logits).abs().max())
    print(f"oracle vs reference: latents rel-L2 {dx:.2e}, logits max abs {dl:.2e}")
    assert dx < 1e-6 and dl < 1e-6
    np.savez_compressed(os.path.join(HERE, "ae_det.npz"), pc=pc.numpy(), queries=q.numpy(),
                        latents_rows=x[0, ::8].numpy(), logits=logits[0, :, 0].numpy(),
                        fps_idx=orc.fps_indices(pc, 512)[0].numpy(), state_hash=np.array(sd_hash(sd)),
                        n_keys=np.int64(len(sd)))
    print("wrote ae_det.npz", len(sd), "state tensors, hash", sd_hash(sd)[:16])


if __name__ == "__main__":
    main()
