"""Dev check (B200): the four-rows-per-warp boundary kernel must give BIT-IDENTICAL sampler output to the one-row form.
Runs a 12-frame, 3-step sampler in two processes (RALD_B200_BOUNDARY_R4 = 1 / 0) and compares SHA-256 of the latents."""
import hashlib, os, subprocess, sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, "."); sys.path.insert(0, "tests")
    import torch
    from helpers import build_denoiser
    from rald_b200 import synth
    net = build_denoiser(device="cuda")
    tok = torch.randn(12, 64, 512, generator=torch.Generator().manual_seed(1)).cuda()
    lat = synth.unit_latents(range(12)).cuda()
    x = net.sample_from_latents(lat, tok, num_steps=3)
    print("HASH", hashlib.sha256(x.cpu().numpy().tobytes()).hexdigest(), float(x.abs().mean()))
else:
    out = []
    for v in ("1", "0"):
        env = dict(os.environ, RALD_B200_BOUNDARY_R4=v)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("HASH")]
        print(f"R4={v}:", line[0] if line else r.stderr[-500:])
        out.append(line[0] if line else None)
    ok = out[0] is not None and out[0] == out[1]
    print("BIT-IDENTICAL" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
