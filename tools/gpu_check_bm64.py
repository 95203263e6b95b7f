"""Dev check (B200): the 64-row-tile form of the GEMM (batch-1 regime) must be BIT-IDENTICAL to the 128-row tiles
(explicit bn hint -> 128-row path) for every small-M shape it serves. python tools/gpu_check_bm64.py"""
import sys
sys.path.insert(0, ".")
import torch
from rald_b200 import _lib

torch.manual_seed(0)
dev = "cuda"
ok_all = True
for (M, N, K, mode, inplace) in [(512, 512, 512, 1, True), (512, 512, 2048, 1, True), (512, 512, 512, 0, False),
                                 (500, 512, 512, 1, True), (64, 512, 512, 0, False), (384, 256, 1024, 1, False),
                                 (1024, 512, 512, 1, True), (70, 64, 64, 0, False)]:
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    h0 = torch.randn(M, N, device=dev)
    outs = []
    for bn in (0, 32 if mode == 1 else 64):
        if mode == 1:
            out = h0.clone()
            r = out if inplace else h0
        else:
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            r = None
        _lib.call("rald_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, b.data_ptr(), _lib.ptr(r), N,
                  M, N, K, mode, bn, _lib.cur_stream())
        torch.cuda.synchronize()
        outs.append(out.clone())
    ref = A.float() @ W.float().t() + b + (h0 if mode == 1 else 0)
    err = float((outs[0].float() - ref).norm() / ref.norm())
    same = torch.equal(outs[0], outs[1])
    ok = same and err < (1e-5 if mode == 1 else 6e-3)
    ok_all &= ok
    print(f"M={M} N={N} K={K} mode={mode} inplace={inplace}: auto == hinted bitwise: {same}  rel vs fp32 {err:.2e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
print("ALL OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
