"""Isolated timing of the residual-GEMM shape with different epilogues (CUDA events, back-to-back, L2-warm)."""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from rald_b200 import _lib
dev = "cuda"
_lib.lib()

def bench(M, N, K, mode, inplace, bn=0, iters=50):
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N if mode != 2 else N // 2, device=dev, dtype=torch.float32 if mode == 1 else torch.bfloat16)
    r = out if inplace else None
    args = (A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), out.shape[1], b.data_ptr(), _lib.ptr(r), N if r is not None else 0,
            M, N, K, mode, bn, _lib.cur_stream())
    for _ in range(5): _lib.call("rald_gemm_bf16", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): _lib.call("rald_gemm_bf16", *args)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"M={M} N={N} K={K} mode={mode} inplace={inplace} bn={bn}: {us:.1f} us {2.0*M*N*K/us/1e6:.0f} TFLOP/s", flush=True)

for M in (16384, 32768):
    for bn in (0, 128, 256):
        bench(M, 512, 512, 1, True, bn)
        bench(M, 512, 512, 1, False, bn)
        bench(M, 512, 512, 0, False, bn)
    bench(M, 512, 2048, 1, True)
    bench(M, 512, 2048, 1, False)
    bench(M, 1536, 512, 0, False)
