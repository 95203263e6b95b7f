"""GPU diagnostic: where the common-mode error of the random-init occupancy field comes from. Splits the decode path
into sampler / latent stack / fold + query kernel by swapping each stage with the fp32 CPU oracle.
    python tools/gpu_chamfer_probe.py            (on a B200; prints a table)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_ae, build_denoiser, cpu_state_dict  # noqa: E402
from oracle import rald_oracle as orc  # noqa: E402
from rald_b200 import synth  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "e2e.npz"))
ref = torch.from_numpy(g["logits"][0])
z_ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "sampler_trace.npz"))["trace"][-1])[None]
q = synth.query_points(1, 32768, seed=99)
field = float(ref.std())


def show(name, lg):
    err = lg.double() - ref.double()
    print(f"{name:70s} common-mode {float(err.mean()):+.3e} ({float(err.mean()) / field:+.3f} sigma)  residual "
          f"{float((err - err.mean()).std()):.2e}", flush=True)


with torch.no_grad():
    net = build_denoiser(device="cuda")
    z_ours = net.sample_from_latents(synth.unit_latents([0]).cuda(), synth.radar_cube(1, seed=1024).cuda()).cpu()
    print("sampler: final latents rel-L2 vs reference", orc.rel_l2(z_ours, z_ref))
    vae = build_ae("kl_d512_m512_l32_mix", device="cuda")
    sd = cpu_state_dict(vae)
    x_ref = orc.ae_latent_stack(sd, z_ref)
    show("oracle decode of reference latents (sanity: 0)", orc.ae_query(sd, x_ref, q)[0, :, 0])
    show("oracle decode of OUR latents (sampler alone)", orc.ae_decode(sd, z_ours, q)[0, :, 0])
    for precise in ("1", "0"):
        os.environ["RALD_B200_AE_PRECISE"] = precise
        rt = vae._runtime()
        x_gpu = rt.latent_stack(z_ref.cuda()).view(1, 512, 512).cpu()
        print(f"[precise={precise}] stack rel-L2 vs oracle {orc.rel_l2(x_gpu, x_ref):.2e}")
        show(f"[precise={precise}] GPU stack (ref latents) + oracle query", orc.ae_query(sd, x_gpu, q)[0, :, 0])
        rt.clear_cache()
        show(f"[precise={precise}] GPU stack + GPU fold/query (ref latents)", vae.decode(z_ref.cuda(), q.cuda())[0, :, 0].cpu())
        rt.clear_cache()
        show(f"[precise={precise}] full GPU path (our latents)", vae.decode(z_ours.cuda(), q.cuda())[0, :, 0].cpu())
    # fold / query kernel alone: oracle stack output fed to the GPU fold + query kernel
    os.environ["RALD_B200_AE_PRECISE"] = "1"
    rt = vae._runtime()
    rt.ensure_packed()
    ctx = rt._fold_context(x_ref.view(512, 512).cuda().contiguous(), 1)
    show("oracle stack + GPU fold/query kernel", rt.query(ctx, q.cuda())[0].cpu())
