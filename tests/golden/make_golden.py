"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference, this
container only) on seeded synthetic inputs, and checks the CPU oracle (oracle/rald_oracle.py) against them.

    python tests/golden/make_golden.py            # ~3 min on 8 cores

Weights are never stored: every model is rebuilt from ``torch.manual_seed(1024)`` (the drop-in modules consume
the RNG exactly like the reference; a SHA-256 of the state dict is stored and re-checked by the tests) with
``model.proj_out`` re-initialised, because the reference zero-initialises it (models_radar_generation.py:198)
which would make the network output identically zero.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import ref_import  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402

SEED = 1024


def sd_hash(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_denoiser(factory, name, configs):
    torch.manual_seed(SEED)
    net = factory[name](configs=configs).eval()
    net.model.proj_out.reset_parameters()  # draws from the global RNG right after construction
    return net


def report(tag, a, b, tol):
    e = orc.rel_l2(a, b)
    print(f"  oracle vs reference {tag}: rel-L2 {e:.3e} (tol {tol:g})", flush=True)
    assert e <= tol, f"oracle deviates from the reference at {tag}: {e}"
    return e


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    m_ae, m_gen, m_enc = ref_import.import_reference()
    cfg = ref_import.load_generation_config()
    meta = {"seed": SEED, "torch": torch.__version__, "hashes": {}, "oracle_vs_reference": {}}
    t0 = time.time()

    # ------------------------------------------------------------------ denoiser + radar encoder
    net = build_denoiser(m_gen.__dict__, cfg.ar_model.name, cfg.ar_model.configs)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    meta["hashes"]["denoiser"] = sd_hash(sd)
    cube = synth.radar_cube(1, seed=SEED)
    cube_sparse = synth.radar_cube(1, seed=SEED + 1, sparse=True)
    out = {}
    for tag, cb in (("dense", cube), ("sparse", cube_sparse)):
        enc_ref = net.radar_enc(cb[..., 0:1].permute(0, 4, 1, 2, 3))
        tok_ref = net.process_radar_cond(cb)
        enc_orc = orc.radar_encoder(sd, cb[..., 0:1].permute(0, 4, 1, 2, 3))
        tok_orc = orc.process_radar_cond(sd, cb)
        meta["oracle_vs_reference"][f"radar_enc_{tag}"] = report(f"radar_enc[{tag}]", enc_orc, enc_ref, 1e-5)
        meta["oracle_vs_reference"][f"tokens_{tag}"] = report(f"tokens[{tag}]", tok_orc, tok_ref, 1e-5)
        out[f"enc_{tag}"] = enc_ref.numpy()
        out[f"tokens_{tag}"] = tok_ref.numpy()
    np.savez_compressed(os.path.join(HERE, "radar_cond.npz"), **out)
    print(f"[{time.time()-t0:.0f}s] radar conditioning done", flush=True)

    # one network evaluation at three noise levels (shared sigma) and with per-sample sigma (training-style call)
    lat = synth.unit_latents([0, 1])
    cube2 = synth.radar_cube(2, seed=SEED)
    tok2 = net.process_radar_cond(cube2)
    evals = {"tokens2": tok2.numpy()}
    for sg in (80.0, 1.5, 0.002):
        x = lat * sg
        d_ref = net(x, torch.tensor(sg), cube2, "radar")
        d_orc = orc.edm_precond(sd, x, torch.tensor(sg), tok2)
        meta["oracle_vs_reference"][f"denoised_sigma{sg}"] = report(f"denoised sigma={sg}", d_orc, d_ref, 2e-5)
        evals[f"denoised_{sg}"] = d_ref.numpy()
    sg_ps = torch.tensor([3.0, 0.2]).reshape(2, 1, 1)
    d_ref = net(lat * sg_ps, sg_ps, cube2, "radar")
    d_orc = orc.edm_precond(sd, lat * sg_ps, sg_ps, tok2)
    meta["oracle_vs_reference"]["denoised_per_sample"] = report("denoised per-sample sigma", d_orc, d_ref, 2e-5)
    evals["denoised_per_sample"] = d_ref.numpy()
    np.savez_compressed(os.path.join(HERE, "denoiser_eval.npz"), **evals)
    print(f"[{time.time()-t0:.0f}s] single evaluations done", flush=True)

    # full 18-step sampler, B = 1, through the reference's own edm_sampler (35 net evals, encoder re-run each time)
    class Recorder:
        """Wraps the reference net to record the un-scaled input of every evaluation."""
        def __init__(self, inner):
            self.inner, self.calls = inner, []
            self.sigma_min, self.sigma_max = inner.sigma_min, inner.sigma_max
        def round_sigma(self, s):
            return self.inner.round_sigma(s)
        def __call__(self, x, sigma, labels, cond_type):
            self.calls.append(x.clone())
            return self.inner(x, sigma, labels, cond_type)
    rec = Recorder(net)
    lat1 = synth.unit_latents([0])
    x_final = m_gen.edm_sampler(rec, lat1, cube, "radar")
    assert len(rec.calls) == 35
    # x_next after step i (i < 17) is the input of the first evaluation of step i+1 = call 2*(i+1)
    trace_ref = torch.stack([rec.calls[2 * (i + 1)] for i in range(17)] + [x_final])[:, 0]
    print(f"[{time.time()-t0:.0f}s] reference sampler done", flush=True)
    tr = []
    x_orc = orc.edm_sample(sd, lat1, torch.from_numpy(out["tokens_dense"]), trace=tr)
    trace_orc = torch.stack(tr)[:, 0]
    meta["oracle_vs_reference"]["sampler_final"] = report("sampler final latents", x_orc, x_final, 1e-4)
    meta["oracle_vs_reference"]["sampler_trace"] = report("sampler per-step latents", trace_orc, trace_ref, 1e-4)
    np.savez_compressed(os.path.join(HERE, "sampler_trace.npz"), trace=trace_ref.numpy())
    print(f"[{time.time()-t0:.0f}s] oracle sampler done", flush=True)
    del net, sd

    # ------------------------------------------------------------------ autoencoders
    ae_out = {}
    for name, qtype in (("kl_d512_m512_l32_mix", "mix"), ("kl_d512_m512_l32", "point")):
        torch.manual_seed(SEED)
        ae = m_ae.__dict__[name](N=10000).eval()
        sd = {k: v.detach() for k, v in ae.state_dict().items()}
        meta["hashes"][name] = sd_hash(sd)
        pc = synth.lidar_points(1, 10000, seed=SEED) if qtype == "mix" else synth.frustum_points(1, 10000, seed=SEED)
        torch.manual_seed(7)
        kl_ref, z_ref = ae.encode(pc)
        mean, logvar = orc.ae_encode_stats(sd, pc, qtype)
        torch.manual_seed(7)
        noise = torch.randn(mean.shape)
        kl_orc, z_orc = orc.ae_posterior(mean, logvar, noise)
        meta["oracle_vs_reference"][f"{name}_z"] = report(f"{name} z", z_orc, z_ref, 1e-5)
        meta["oracle_vs_reference"][f"{name}_kl"] = report(f"{name} kl", kl_orc, kl_ref, 1e-5)
        ae_out[f"{qtype}_mean"] = mean.numpy()
        ae_out[f"{qtype}_logvar"] = logvar.numpy()
        ae_out[f"{qtype}_kl"] = kl_ref.numpy()
        if qtype == "point":
            ae_out["point_fps_idx"] = orc.fps_indices(pc, 512).numpy()
        if qtype == "mix":
            z = synth.posterior_noise(1, seed=11)  # a latent set to decode
            q = synth.query_points(1, 8192)
            logits_ref = ae.decode(z, q)
            stack_orc = orc.ae_latent_stack(sd, z)
            logits_orc = orc.ae_query(sd, stack_orc, q)
            meta["oracle_vs_reference"]["decode_logits"] = report("decode logits", logits_orc, logits_ref, 1e-4)
            ae_out["decode_logits"] = logits_ref.numpy()
            ae_out["decode_stack_rows"] = stack_orc[0, :64].numpy()  # first 64 latent rows of the stack output
        print(f"[{time.time()-t0:.0f}s] {name} done", flush=True)
    np.savez_compressed(os.path.join(HERE, "ae.npz"), **ae_out)
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
