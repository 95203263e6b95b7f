"""Dev check (run on a B200 through gpurun): tcgen05 GEMM numerics vs torch fp32 and raw throughput."""
import sys, time
sys.path.insert(0, ".")
import torch
from rald_b200 import _lib

torch.manual_seed(0)
dev = "cuda"


def run(M, N, K, mode, bn=0, bias=True, resid=False, inplace=False):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    ref = A.float() @ W.float().t()
    if b is not None:
        ref = ref + b
    if mode == 2:
        # packed layout: each 32 group = 16 value + 16 gate
        g = ref.view(M, N // 32, 2, 16)
        ref = (g[:, :, 0] * torch.nn.functional.gelu(g[:, :, 1])).reshape(M, N // 2)
        out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    elif mode == 1:
        if r is not None:
            ref = ref + r
        out = r.clone() if (inplace and r is not None) else torch.empty(M, N, device=dev, dtype=torch.float32)
        if inplace and r is not None:
            r = out   # h += A W^T + b : TMA reduce-add epilogue
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _lib.call("rald_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), out.shape[1],
              _lib.ptr(b), _lib.ptr(r), N, M, N, K, mode, bn, _lib.cur_stream())
    torch.cuda.synchronize()
    err = (out.float() - ref).norm() / ref.norm()
    mx = (out.float() - ref).abs().max()
    ok = err < (1e-5 if mode == 1 else 6e-3)
    print(f"gemm M={M} N={N} K={K} mode={mode} bn={bn} bias={bias} resid={resid} inplace={inplace}: rel={err:.3e} max={mx:.3e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    return ok


def bench(M, N, K, mode, bn=0, iters=20):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N if mode != 2 else N // 2, device=dev, dtype=torch.float32 if mode == 1 else torch.bfloat16)
    r = out if mode == 1 else None   # in-place residual, as every residual GEMM of the denoiser / AE stack
    args = (A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), out.shape[1], _lib.ptr(b), _lib.ptr(r), N, M, N, K,
            mode, bn, _lib.cur_stream())
    for _ in range(3):
        _lib.call("rald_gemm_bf16", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.call("rald_gemm_bf16", *args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # torch reference
    for _ in range(3):
        torch.matmul(A, W.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, W.t())
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    print(f"bench M={M} N={N} K={K} mode={mode} bn={bn}: {ms*1e3:.1f} us {tf:.0f} TFLOP/s | cuBLAS bf16 "
          f"{ms_t*1e3:.1f} us {2.0*M*N*K/ms_t/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", _lib.lib().rald_abi_version(), flush=True)
    ok = True
    ok &= run(128, 128, 64, 1, bn=128, bias=False)
    ok &= run(128, 128, 64, 0, bn=128, bias=False)
    ok &= run(256, 256, 512, 1, bn=128)
    ok &= run(512, 1536, 512, 0)
    ok &= run(512, 1536, 512, 0, bn=256)
    ok &= run(512, 512, 512, 1, resid=True)
    for bn_ in (0, 32, 64, 128, 256):
        ok &= run(512, 512, 512, 1, resid=True, inplace=True, bn=bn_)
        ok &= run(1000, 512, 2048, 1, resid=True, inplace=True, bn=bn_)   # ragged M through the TMA reduce
        ok &= run(300, 512, 512, 1, bn=bn_)                                # plain fp32 TMA store
    for bn_ in (64, 128, 256):
        ok &= run(1000, 1536, 512, 0, bn=bn_)
    for bn_ in (128, 256):
        ok &= run(1000, 4096, 512, 2, bn=bn_)
    ok &= run(512, 10016, 64, 1)         # N multiple of 32 only (long-context scores)
    ok &= run(512, 10016, 512, 0)        # bf16 output, N % 64 != 0 -> generic epilogue
    ok &= run(40000, 512, 512, 1, resid=True, inplace=True)
    ok &= run(512, 512, 2048, 1, resid=True, bn=256)
    ok &= run(512, 4096, 512, 2)
    ok &= run(512, 4096, 512, 2, bn=256)
    ok &= run(4096, 4096, 512, 2, bn=256)
    ok &= run(1000, 512, 64, 0)        # ragged M
    ok &= run(10000, 1024, 512, 0)     # AE encoder K/V projection, ragged M
    ok &= run(512, 512, 32, 1, bias=False)  # K < 64 (TMA zero fill)
    ok &= run(512, 64, 512, 1, bn=64)
    ok &= run(512, 32, 512, 1, bn=32)
    ok &= run(64, 24576, 512, 0)       # hoisted ctx K/V
    ok &= run(32768, 1536, 512, 0, bn=256)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    for (M, N, K, mode, bn) in [(32768, 1536, 512, 0, 256), (32768, 1536, 512, 0, 128), (32768, 4096, 512, 2, 256),
                                (32768, 512, 2048, 1, 256), (32768, 512, 2048, 1, 128), (32768, 512, 512, 1, 256),
                                (32768, 512, 512, 1, 128),
                                (4096, 1536, 512, 0, 256), (4096, 1536, 512, 0, 128), (4096, 4096, 512, 2, 256),
                                (4096, 512, 2048, 1, 128), (4096, 512, 2048, 1, 64),
                                (512, 1536, 512, 0, 128), (512, 1536, 512, 0, 64), (512, 4096, 512, 2, 128),
                                (512, 512, 2048, 1, 64), (512, 512, 2048, 1, 32),
                                (8192, 8192, 8192, 0, 256)]:
        bench(M, N, K, mode, bn)
    sys.exit(0 if ok else 1)
