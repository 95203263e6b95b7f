"""Times the fused cross-attention kernel (rald_xattn_fused) alone at the bench shape: python tools/gpu_time_xattn.py [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rald_b200 import _lib

F = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
dev, bf, dim, M, depth = "cuda:0", torch.bfloat16, 512, 512, 4
torch.manual_seed(0)
kp = (torch.randn(depth, 8, F, 64, dim, device=dev) * 0.05).to(bf)
vt = (torch.randn(depth, 8, dim, F * 64, device=dev) * 0.5).to(torch.float16)
xn = torch.randn(F * M, dim, device=dev).to(bf)
h = torch.zeros(F * M, dim, device=dev)
bias = torch.zeros(dim, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
st = _lib.cur_stream()
def run(n):
    _lib.call("rald_xattn_fused", xn.data_ptr(), kp[n % depth].data_ptr(), vt[n % depth].data_ptr(), bias.data_ptr(),
              h.data_ptr(), F, M, 0, F, st)
for i in range(5): run(i)
torch.cuda.synchronize()
ts = []
for i in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(i); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
flops = 4.0 * F * M * dim * 512
print(f"frames={F}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  -> {flops / (ts[len(ts)//2] * 1e-6) / 1e12:.0f} TFLOP/s (L2 flushed between launches)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): run(i)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back: {e0.elapsed_time(e1) * 10:.1f} us per launch")

if "--phases" in sys.argv:
    dbg = torch.zeros(64, device=dev, dtype=torch.int64)
    _lib.call("rald_xattn_debug_buffer", dbg.data_ptr())
    run(0); torch.cuda.synchronize()
    _lib.call("rald_xattn_debug_buffer", 0)
    d = dbg.cpu().view(4, 16)
    t0 = int(d[0, 0])
    names = ["mma:start", "mma:S issued", "mma:P0 ready", "mma:P1 ready", "mma:O0 issued", "mma:O1 issued", "mma:O2 issued",
             "mma:O3 issued", "sm:S ready", "sm:P written", "sm:O0 ready", "sm:O1 ready", "sm:O2 ready", "sm:O3 ready",
             "sm:tile stored"]
    for t in range(2):
        print(f"tile {t}: " + "  ".join(f"{n}={(int(d[t, i]) - t0) / 1e3:.2f}" for i, n in enumerate(names) if int(d[t, i])))
