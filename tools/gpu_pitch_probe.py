"""Does the row pitch of the K-major operands matter for the split-K GEMM? (conv wgrad: pitch 5.5 MB, 4.2 TB/s of SM
ingest; denoiser wgrad: pitch 64 KB, 11.5 TB/s)"""
import sys
import torch
sys.path.insert(0, ".")
from rald_b200 import _lib
DEV, BF = "cuda:0", torch.bfloat16
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M, N = 512, 2048
for K, pitch in ((32768, 32768), (32768, 2752512), (262144, 262144), (262144, 2752512), (2752512, 2752512)):
    A = torch.randn(M, pitch, device=DEV).to(BF) if pitch * M < 3e9 else None
    A = torch.empty(M, pitch, device=DEV, dtype=BF).normal_()
    Bm = torch.empty(N, pitch, device=DEV, dtype=BF).normal_()
    out = torch.zeros(M, N, device=DEV)
    ms = timed(lambda: _lib.call("rald_gemm_bf16_accum", A.data_ptr(), pitch, Bm.data_ptr(), pitch, out.data_ptr(), N, M, N, K, _lib.cur_stream()))
    print(f"K {K} pitch {pitch}: {ms:.3f} ms, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    del A, Bm
