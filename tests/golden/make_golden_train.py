"""Adds tests/golden/train_grads.npz: ONE training step's loss and parameter gradients from the UNMODIFIED reference
(EDMLoss arithmetic, model/models_radar_generation.py:277-295, through EDMPrecond.forward under autograd, as
engine_generation.py:89-110 runs it) on the default denoiser in .train() mode, 2 frames, with the radar encoder frozen
(the reference's frozen-encoder option, engine_generation.py:86-87) and the EDMLoss draws (sigma, noise) recorded so
the GPU test can inject them. Also checks that autograd through the CPU oracle reproduces the reference's gradients.

Stored: loss, D(y + n; sigma), sigma, noise; for every trainable parameter its gradient norm; full gradients of the
vectors and small matrices; a fixed 2048-element sample (seeded indices) of every large matrix.

    python tests/golden/make_golden_train.py        # ~2 min, needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))
import ref_import  # noqa: E402
from helpers import grad_sample_index  # noqa: E402
from make_golden import SEED, build_denoiser  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402

SAMPLE = 2048
FULL_BELOW = 70000      # parameters with fewer elements are stored whole


def sample_index(name: str, numel: int) -> np.ndarray:
    return grad_sample_index(name, numel, SAMPLE)


def main():
    torch.set_num_threads(os.cpu_count())
    _, m_gen, _ = ref_import.import_reference()
    cfg = ref_import.load_generation_config()
    net = build_denoiser(m_gen.__dict__, cfg.ar_model.name, cfg.ar_model.configs)
    net.train()
    net.radar_enc.requires_grad_(False)
    B = 2
    cube = synth.radar_cube(B, seed=SEED)
    y = synth.unit_latents([5, 6]) * 0.7
    g = torch.Generator("cpu").manual_seed(31)
    rnd = torch.randn([B, 1, 1], generator=g)
    sigma = (rnd * 1.2 - 1.2).exp()
    noise = torch.randn(y.shape, generator=g)
    weight = (sigma ** 2 + 1.0) / (sigma * 1.0) ** 2

    D = net(y + noise * sigma, sigma, cube, "radar")
    loss = (weight * (D - y) ** 2).mean()
    loss.backward()
    out = {"loss": np.float64(loss.item()), "D": D.detach().numpy(), "sigma": sigma.numpy(), "noise": noise.numpy(),
           "y": y.numpy()}
    ref_grads = {}
    names = []
    for name, p in net.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        gr = p.grad.detach()
        ref_grads[name] = gr
        names.append(name)
        out["norm/" + name] = np.float64(gr.double().norm().item())
        if gr.numel() <= FULL_BELOW:
            out["full/" + name] = gr.numpy()
        else:
            out["sample/" + name] = gr.reshape(-1)[torch.from_numpy(sample_index(name, gr.numel()))].numpy()
    out["names"] = np.array(names)
    print(f"loss {loss.item():.6f}; {len(names)} trainable tensors", flush=True)

    # ---- the oracle under autograd reproduces these gradients (tokens from the frozen encoder) ----
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        feat_tokens_ref = net.process_radar_cond(cube)
    for k in list(sd.keys()):
        if not k.startswith("radar_enc."):
            sd[k].requires_grad_(True)
    tok = orc.process_radar_cond(sd, cube)
    assert orc.rel_l2(tok.detach(), feat_tokens_ref) < 1e-5
    D_o = orc.edm_precond(sd, y + noise * sigma, sigma, tok)
    loss_o = (weight * (D_o - y) ** 2).mean()
    loss_o.backward()
    worst = 0.0
    for name in names:
        e = orc.rel_l2(sd[name].grad, ref_grads[name])
        worst = max(worst, e)
        assert e < 2e-4, (name, e)
    print(f"oracle autograd vs reference: loss {abs(loss_o.item() - loss.item()):.2e}, worst gradient rel-L2 {worst:.2e}")
    np.savez_compressed(os.path.join(HERE, "train_grads.npz"), **out)

    # ---- the same step with the radar encoder TRAINABLE (unfreeze_radar_enc: true, the shipped configuration): the
    # denoiser's gradients are unchanged (asserted), the encoder's 144 tensors are stored in train_grads_enc.npz ----
    net.zero_grad(set_to_none=True)
    net.radar_enc.requires_grad_(True)
    D2 = net(y + noise * sigma, sigma, cube, "radar")
    loss2 = (weight * (D2 - y) ** 2).mean()
    loss2.backward()
    assert abs(loss2.item() - loss.item()) < 1e-9
    out2 = {"loss": np.float64(loss2.item())}
    enc_names = []
    for name, p in net.named_parameters():
        gr = p.grad.detach()
        if not name.startswith("radar_enc."):
            assert orc.rel_l2(gr, ref_grads[name]) < 1e-6, name
            continue
        enc_names.append(name)
        out2["norm/" + name] = np.float64(gr.double().norm().item())
        if gr.numel() <= FULL_BELOW:
            out2["full/" + name] = gr.numpy()
        else:
            out2["sample/" + name] = gr.reshape(-1)[torch.from_numpy(sample_index(name, gr.numel()))].numpy()
    out2["names"] = np.array(enc_names)
    print(f"trainable encoder: {len(enc_names)} tensors", flush=True)
    np.savez_compressed(os.path.join(HERE, "train_grads_enc.npz"), **out2)


if __name__ == "__main__":
    main()
