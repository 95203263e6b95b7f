"""Phase breakdown of the tcgen05 GEMM from in-kernel %globaltimer stamps (rald_gemm_debug_buffer), plus CUDA-event
timing of back-to-back launches. python tools/gemm_phases.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
_lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(M, N, K, mode, bn, reps=20):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N // 2 if mode == 2 else N, device=dev, dtype=torch.float32 if mode == 1 else torch.bfloat16)
    resid = out if mode == 1 else None
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    st = _lib.cur_stream()

    def call():
        _lib.call("rald_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), out.shape[1], bias.data_ptr(),
                  _lib.ptr(resid), out.shape[1] if resid is not None else 0, M, N, K, mode, bn, st)
    call()
    torch.cuda.synchronize()
    _lib.lib().rald_gemm_debug_buffer(dbg.data_ptr())
    acc = None
    for _ in range(reps):
        flush.zero_()
        dbg.zero_()
        call()
        torch.cuda.synchronize()
        d8 = dbg.view(148, 8).cpu()
        ring = d8[d8[:, 0] > 0]
        ring_issued = (ring[:, 7] - ring[:, 0].min()).float().mean() if (ring[:, 7] > 0).all() else torch.tensor(float("nan"))
        d = d8[:, :7]
        d = d[d[:, 0] > 0]
        t0 = d[:, 0].min()
        rel = (d - t0).float()
        row = torch.cat([rel.mean(0), rel[:, 6].max()[None], torch.tensor([float(d.shape[0])]), ring_issued[None]])
        acc = row if acc is None else acc + row
    _lib.lib().rald_gemm_debug_buffer(0)
    acc /= reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        call()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 50 * 1000
    names = ["entry", "setup", "1st-ops", "mma-issued", "acc-ready", "epi-done", "exit"]
    print(f"M={M} N={N} K={K} mode={mode} bn={bn} ctas={int(acc[8])}: " +
          " ".join(f"{n}={acc[i] / 1000:.2f}" for i, n in enumerate(names)) +
          f" | last-exit={acc[7] / 1000:.2f} us | ring-issued={acc[9] / 1000:.2f} | warm back-to-back {warm:.2f} us")


if __name__ != "__main__":
    SHAPES = []
else:
    SHAPES = None
for (M, N, K, mode) in SHAPES if SHAPES is not None else [(512, 1536, 512, 0), (512, 512, 512, 1), (512, 4096, 512, 2), (512, 512, 2048, 1),
                        (4096, 1536, 512, 0), (4096, 512, 2048, 1)]:
    for bn in ((32, 64, 128) if M == 512 else (128, 256)):
        if mode == 2 and bn < 64:
            continue
        run(M, N, K, mode, bn)
