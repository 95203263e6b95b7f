import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from rald_b200 import _lib
DEV = "cuda:0"; BF = torch.bfloat16
s = _lib.cur_stream()
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
M, N, K = 64, 64, 16384
g = torch.Generator().manual_seed(1)
A = torch.randn(M, K, generator=g).to(DEV).to(BF); W = torch.randn(N, K, generator=g).to(DEV).to(BF)
for shift in (0, 64, 1, -1, -101):
    out = torch.zeros(M, N, device=DEV)
    _lib.call("rald_gemm_bf16_accum_shift", A.data_ptr(), K, W.data_ptr(), K, shift, out.data_ptr(), N, M, N, K, s)
    sync(f"gemm shift {shift}")
    Wf = W.float()
    Ws = torch.zeros_like(Wf)
    if shift >= 0: Ws[:, :K - shift] = Wf[:, shift:]
    else: Ws[:, -shift:] = Wf[:, :K + shift]
    ref = A.float() @ Ws.t()
    print("   rel", float((out - ref).norm() / ref.norm()), flush=True)
