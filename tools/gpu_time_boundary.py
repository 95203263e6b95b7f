"""Dev timing (B200): rald_dit_boundary (Euler mode, next projection on) at 1 / 8 / 64 / 256 frames. Chains of launches
rotating over enough buffer sets to exceed the L2, captured in a CUDA graph (no host launch cost in the figure) and
timed with CUDA events; prints us per launch, the algorithmic HBM bytes (h read + h_next written + the [T, C]
vectors) and the fraction of MEASURED_PEAKS.json's copy bandwidth.   --frames F [--plain]: one size, no graph (ncu)."""
import json, sys
sys.path.insert(0, ".")
import torch
from rald_b200 import _lib

try:
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6544.7)
except OSError:
    peak = 6544.7   # the pool's measured copy bandwidth when the driver-written file is absent
C, rows = 32, 512
g = torch.Generator("cuda").manual_seed(0)
ln_w = torch.ones(512, device="cuda"); ln_b = torch.zeros(512, device="cuda")
w_out_t = torch.randn(512, 32, device="cuda", generator=g) / 22.6
w_in_t = torch.randn(C, 512, device="cuda", generator=g) / 5.6
pack = torch.empty(_lib.boundary_pack_bytes(), device="cuda", dtype=torch.uint8)
_lib.call("rald_dit_boundary_pack", ln_w.data_ptr(), ln_b.data_ptr(), w_out_t.data_ptr(), w_in_t.data_ptr(), C,
          pack.data_ptr(), _lib.cur_stream())
sizes = (1, 8, 64, 256)
plain = "--plain" in sys.argv
if "--frames" in sys.argv:
    sizes = (int(sys.argv[sys.argv.index("--frames") + 1]),)
for frames in sizes:
    T = frames * rows
    nset = max(2, min(24, int(300e6 // (T * 4096)) + 1))
    sets = []
    for _ in range(nset):
        sets.append(dict(h=torch.randn(T, 512, device="cuda", generator=g), hn=torch.empty(T, 512, device="cuda"),
                         x=torch.randn(T, C, device="cuda", generator=g), d=torch.empty(T, C, device="cuda"),
                         xo=torch.empty(T, C, device="cuda")))
    sig = torch.full((frames,), 1.5, device="cuda"); sig_o = torch.full((frames,), 1.1, device="cuda")
    def launch(s, stream):
        _lib.call("rald_dit_boundary", s["h"].data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), w_out_t.data_ptr(),
                  w_in_t.data_ptr(), s["x"].data_ptr(), 0, s["d"].data_ptr(), s["xo"].data_ptr(), s["hn"].data_ptr(),
                  sig.data_ptr(), 1, sig_o.data_ptr(), 1, 1, rows, C, T, 512, 0.5, pack.data_ptr(), stream)
    for i in range(2 * nset):
        launch(sets[i % nset], _lib.cur_stream())
    torch.cuda.synchronize()
    n = 10 * nset
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if plain:
        e0.record()
        for i in range(n):
            launch(sets[i % nset], _lib.cur_stream())
        e1.record(); torch.cuda.synchronize()
        reps = 1
    else:
        st = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            with torch.cuda.graph(graph, stream=st):
                for i in range(n):
                    launch(sets[i % nset], st.cuda_stream)
        graph.replay(); torch.cuda.synchronize()
        reps = 5
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (n * reps)
    byts = T * (2048 + 2048 + C * 16)
    print(f"frames {frames:3d}  rows {T:6d}  {us:8.2f} us / launch   {byts / us / 1e3:8.1f} GB/s = {byts / us / 1e3 / peak:.3f} of {peak:.0f}")
