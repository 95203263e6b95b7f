"""End-to-end north-star criterion on the GPU: radar cube -> encoder -> 18-step sampler -> VecSet decode, against the
reference's own result for the same cube, weights and injected noise.

Point clouds = occupied query points (`logit > 0`, engine_generation.py:285) mapped back with inverse_norm_points +
polar2cartesian as `evaluate` does; Chamfer distance per utils/utils.py:116-142 against the frame's lidar cloud (GT):
    |CD(ours, GT) - CD(ref, GT)| / CD(ref, GT) <= 1 %.

Three fixtures / conventions, all with the SHARED threshold (the literal `logit > 0` on both sides):
  * random init (tests/golden/e2e.npz): the reference's field has mean -0.165 and a spatial standard deviation of
    1.5e-3, nothing above 0. SURVEY.md §7.3: the reference's 95-th percentile logit is subtracted from
    `to_outputs.bias` in the SHARED state_dict. A common-mode error of 0.1 % of the logit moves the occupied
    fraction by a quarter here, so this variant needs the split-weight ("precise") latent stack, which is the default
    (rald_b200/runtime_ae.py: precise_enabled);
  * well-conditioned (tests/golden/e2e_wc.npz, make_golden_e2e_wc.py): the same seeded weights with the decoder
    sharpened in the shared state_dict (to_q x 32, to_outputs x 8, bias at the 95-th percentile) so that the logits
    are O(0.1 - 1) as with trained weights;
  * a probe that isolates the SAMPLER: our bf16 sampler's latents decoded by the fp32 CPU oracle, i.e. the part of the
    deviation that no decoder precision can remove.
The per-side-percentile variant of round 1 (each side thresholded at its own 95-th percentile) is kept as a
secondary check of the spatial field."""
import numpy as np
import pytest
import torch

from helpers import build_ae, build_denoiser, cpu_state_dict
from oracle import rald_oracle as orc
from rald_b200 import postproc, synth

pytestmark = pytest.mark.gpu
PC_RANGE = [0, -90, -20, 15.8, 90, 20]


def _gt_cloud():
    return orc.occupancy_points(np.ones(10000, np.float32), synth.lidar_points(1, 10000, seed=1024)[0].numpy(), PC_RANGE,
                                True, False, True)


def _ref_cloud(ref_logits, q, thr=0.0):
    return orc.occupancy_points(ref_logits - np.float32(thr), q[0].numpy(), PC_RANGE, True, False, True)


def _our_latents():
    net = build_denoiser(device="cuda")
    cube = synth.radar_cube(1, seed=1024).cuda()
    return net.sample_from_latents(synth.unit_latents([0]).cuda(), cube)


def _report(name, c_ours, c_ref, gt):
    cd_ours, cd_ref = orc.chamfer_distance(c_ours, gt), orc.chamfer_distance(c_ref, gt)
    rel = abs(cd_ours - cd_ref) / cd_ref
    print(f"[{name}] occupied: ours {len(c_ours)} ref {len(c_ref)}; Chamfer vs GT: ours {cd_ours:.5f} ref {cd_ref:.5f} "
          f"-> {rel:.3%}; CD(ours, ref) = {orc.chamfer_distance(c_ours, c_ref):.5f}")
    return rel


def _stats(name, ours, ref):
    err = ours - ref
    off, resid, field = float(err.mean()), float((err - err.mean()).std()), float(ref.std())
    print(f"[{name}] logits: common-mode error {off:+.3e}, spatial residual {resid:.3e}, field std {field:.3e}")
    return off, resid, field


def test_random_init_shared_threshold(golden):
    """SURVEY.md §7.3 convention, asserted: the reference's 95-th percentile is subtracted from to_outputs.bias on
    BOTH sides and `logit > 0` is applied literally."""
    g = golden("e2e")
    ref_logits = g["logits"][0].numpy()
    shift = float(g["shift"])
    vae = build_ae("kl_d512_m512_l32_mix", device="cuda")
    with torch.no_grad():
        vae.to_outputs.bias -= shift
    z = _our_latents()
    q = synth.query_points(1, 32768, seed=99)
    logits = vae.decode(z, q.cuda())[..., 0]
    ours = logits[0].cpu().numpy()
    ref_shifted = ref_logits - np.float32(shift)
    off, resid, field = _stats("random init", ours, ref_shifted)
    assert orc.rel_l2(torch.from_numpy(ours + np.float32(shift)), torch.from_numpy(ref_logits)) <= 1e-2
    assert resid <= 0.05 * field
    assert abs(off) <= 0.05 * field, "common-mode error of the occupancy field (is the precise latent stack on?)"
    gt = _gt_cloud()
    pts, cnt, _ = postproc.occupied_points(logits, q.cuda(), 0.0, PC_RANGE, True, False, True)
    c_ours = pts[0, :int(cnt[0])].cpu().numpy()
    rel = _report("random init, shared threshold", c_ours, _ref_cloud(ref_shifted, q), gt)
    assert rel <= 0.01
    # secondary: each side at its own 95-th percentile (the spatial field alone)
    thr = float(np.quantile(ours, 0.95))
    pts, cnt, _ = postproc.occupied_points(logits - thr, q.cuda(), 0.0, PC_RANGE, True, False, True)
    rel_own = _report("random init, own 95-th percentile", pts[0, :int(cnt[0])].cpu().numpy(),
                      _ref_cloud(ref_shifted, q), gt)
    assert rel_own <= 0.01


def test_well_conditioned_literal_threshold(golden):
    """Sharpened decoder in the shared state_dict (logits O(0.1 - 1)): literal `logit > 0` on both sides."""
    g = golden("e2e_wc")
    ref_logits = g["logits"][0].numpy()
    vae = build_ae("kl_d512_m512_l32_mix", device="cuda")
    with torch.no_grad():
        vae.decoder_cross_attn.fn.to_q.weight.mul_(float(g["q_scale"]))
        vae.to_outputs.weight.mul_(float(g["out_scale"]))
        vae.to_outputs.bias.fill_(float(g["bias"]))
    z = _our_latents()
    q = synth.query_points(1, 32768, seed=99)
    logits = vae.decode(z, q.cuda())[..., 0]
    ours = logits[0].cpu().numpy()
    off, resid, field = _stats("well-conditioned", ours, ref_logits)
    assert resid <= 0.05 * field and abs(off) <= 0.05 * field
    flips = int(((ours > 0) != (ref_logits > 0)).sum())
    print(f"[well-conditioned] occupancy flips {flips} of {int((ref_logits > 0).sum())} occupied")
    gt = _gt_cloud()
    pts, cnt, _ = postproc.occupied_points(logits, q.cuda(), 0.0, PC_RANGE, True, False, True)
    rel = _report("well-conditioned, logit > 0", pts[0, :int(cnt[0])].cpu().numpy(), _ref_cloud(ref_logits, q), gt)
    assert rel <= 0.01


def test_sampler_deviation_alone_probe(golden):
    """Our bf16 sampler's latents decoded by the fp32 CPU oracle with the reference's weights and threshold: the share
    of the end-to-end deviation that belongs to the sampler (per-step latents are within 1e-2 rel-L2 of the
    reference's, tests/test_gpu_denoiser.py), whatever the decoder's precision."""
    g = golden("e2e")
    ref_logits = g["logits"][0].numpy()
    shift = float(g["shift"])
    z = _our_latents().float().cpu()
    sd = cpu_state_dict(build_ae("kl_d512_m512_l32_mix", device="cpu"))
    q = synth.query_points(1, 32768, seed=99)
    with torch.no_grad():
        lg = orc.ae_decode(sd, z, q)[0, :, 0].numpy()
    off, resid, field = _stats("sampler alone (fp32 decode of our latents)", lg, ref_logits)
    gt = _gt_cloud()
    rel = _report("sampler alone, shared threshold", _ref_cloud(lg, q, shift), _ref_cloud(ref_logits, q, shift), gt)
    flips = int(((lg > shift) != (ref_logits > shift)).sum())
    print(f"[sampler alone] occupancy flips {flips} of {int((ref_logits > shift).sum())} occupied; Chamfer deviation {rel:.3%}")
    assert abs(off) <= 0.05 * field and rel <= 0.01
