"""Shared builders for the tests: seeded default models (identical to the reference's seeded init) and hashes."""
import hashlib

import torch

from rald_b200 import models_ae, models_radar_generation
from rald_b200.config import DEFAULT_DENOISER_NAME, default_denoiser_configs

SEED = 1024


def sd_hash(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_denoiser(name=DEFAULT_DENOISER_NAME, device="cpu"):
    """Constructor-random weights under seed 1024 with proj_out re-randomised (it is zero-initialised)."""
    torch.manual_seed(SEED)
    net = models_radar_generation.__dict__[name](configs=default_denoiser_configs()).eval()
    net.model.proj_out.reset_parameters()
    return net.to(device)


def build_ae(name="kl_d512_m512_l32_mix", n=10000, device="cpu"):
    torch.manual_seed(SEED)
    return models_ae.__dict__[name](N=n).eval().to(device)


def cpu_state_dict(module):
    return {k: v.detach().float().cpu() for k, v in module.state_dict().items()}


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_sample_index(name: str, numel: int, sample: int = 2048):
    """Indices of the gradient entries tests/golden/train_grads.npz keeps for a large parameter (seeded by its name)."""
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000000007
    g = torch.Generator("cpu").manual_seed(h % (1 << 31))
    return torch.randint(0, numel, (sample,), generator=g).numpy()
