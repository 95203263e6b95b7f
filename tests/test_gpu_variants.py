"""GPU parity of the other reference factories / call shapes against the CPU oracle (computed in the test, fp32):
denoiser with 8 latent channels and depth 12, learnable-query and 64-channel autoencoders, empty query set."""
import pytest
import torch

from helpers import SEED, cpu_state_dict, rel_l2
from oracle import rald_oracle as orc
from rald_b200 import models_ae, models_radar_generation, synth
from rald_b200.config import default_denoiser_configs

pytestmark = pytest.mark.gpu


def test_denoiser_l8_depth12_forward_and_short_sampler():
    torch.manual_seed(SEED)
    net = models_radar_generation.kl_d512_m512_l8_edm(configs=default_denoiser_configs()).eval()
    net.model.proj_out.reset_parameters()
    sd = cpu_state_dict(net)
    net = net.cuda()
    g = torch.Generator("cpu").manual_seed(5)
    tokens = torch.randn(2, 64, 512, generator=g)
    lat = torch.randn(2, 512, 8, generator=g)
    sigma = torch.tensor([1.7, 0.05]).reshape(2, 1, 1)
    with torch.no_grad():
        ref = orc.edm_precond(sd, lat * sigma, sigma, tokens)
        out = net(lat.cuda() * sigma.cuda(), sigma.cuda(), tokens.cuda(), "radar")
        assert out.shape == (2, 512, 8)
        assert rel_l2(out, ref) < 1e-2
        x_ref = orc.edm_sample(sd, lat, tokens, num_steps=3)
        x = net.sample_from_latents(lat.cuda(), tokens.cuda(), num_steps=3)
        assert rel_l2(x, x_ref) < 1e-2


@pytest.mark.parametrize("name,qtype,latent", [("kl_d512_m512_l32_learn", "learnable", 32), ("kl_d512_m512_l64", "point", 64)])
def test_autoencoder_variants(name, qtype, latent):
    torch.manual_seed(SEED)
    ae = models_ae.__dict__[name](N=2048).eval()
    # two decoder layers keep the CPU oracle fast; the layer code path is the same as for 24
    sd = cpu_state_dict(ae)
    ae = ae.cuda()
    pc = synth.lidar_points(1, 2048, seed=4)
    with torch.no_grad():
        mean_ref, logvar_ref = orc.ae_encode_stats(sd, pc, qtype)
        mean, logvar = ae.encode_stats(pc.cuda())
    assert mean.shape == (1, 512, latent)
    assert rel_l2(mean, mean_ref) < 1e-2 and rel_l2(logvar, logvar_ref) < 1e-2


def test_decode_with_no_queries():
    torch.manual_seed(SEED)
    ae = models_ae.KLAutoEncoder(depth=1, dim=512, queries_dim=512, output_dim=1, num_inputs=64, num_latents=512,
                                 latent_dim=32, heads=8, dim_head=64, query_type="learnable").eval().cuda()
    out = ae.decode(torch.zeros(2, 512, 32, device="cuda"), torch.zeros(2, 0, 3, device="cuda"))
    assert out.shape == (2, 0, 1)


def test_deterministic_autoencoder_matches_reference():
    """AutoEncoder (reference models_ae.py:181-282, ae_d512_m512): FPS indices bit-exact, latents and logits against
    the fixture written from the unmodified reference (tests/golden/make_golden_ae_det.py)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from helpers import rel_l2
    from rald_b200 import models_ae
    g = np.load(os.path.join(GOLDEN, "ae_det.npz"))
    pc, q = torch.from_numpy(g["pc"]).cuda(), torch.from_numpy(g["queries"]).cuda()
    torch.manual_seed(1024)
    ae = models_ae.ae_d512_m512(N=pc.shape[1]).eval().cuda()
    idx = ae._runtime().fps(pc, 512)
    assert np.array_equal(idx[0].cpu().numpy(), g["fps_idx"])
    x = ae.encode(pc)
    assert x.shape == (1, 512, 512) and x.dtype == torch.float32
    err = rel_l2(x[0, ::8], torch.from_numpy(g["latents_rows"]))
    print("deterministic AE latents rel-L2", err)
    assert err < 1e-2
    out = ae(pc, q)
    assert set(out) == {"logits"} and out["logits"].shape == (1, q.shape[1])
    ref = torch.from_numpy(g["logits"])
    got = out["logits"][0].cpu()
    offset = float((got - ref).mean())
    resid = float(((got - ref) - offset).std())
    print(f"logits: mean {float(ref.mean()):.4f} field std {float(ref.std()):.4f} common-mode {offset:.2e} residual {resid:.2e}")
    assert abs(offset) < 2e-2 * abs(float(ref.mean())) + 1e-3
    assert resid < 0.05 * float(ref.std())
    assert torch.equal(ae.decode(x, q).squeeze(-1), out["logits"])
