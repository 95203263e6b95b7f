"""Times the level-0 conv weight gradient (8 frames) and a denoiser wgrad GEMM (CUDA events)."""
import sys
import torch
import torch.nn as nn
sys.path.insert(0, ".")
from rald_b200 import _lib
from rald_b200.runtime_encoder_train import EncoderTrainRuntime
DEV, BF = "cuda:0", torch.bfloat16
rt = EncoderTrainRuntime.__new__(EncoderTrainRuntime)
rt.dev = torch.device(DEV); rt.groups, rt.eps = 32, 1e-6
B, dims, c = 8, (128, 64, 32), 64
V = dims[0] * dims[1] * dims[2]
x16 = torch.randn(B, V, c, device=DEV).to(BF)
dy = torch.randn(B, V, c, device=DEV)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("conv wgrad (pad/transposes + GEMM) ms", timed(lambda: rt._conv3_wgrad(dy, x16, c, c, dims, 1)))
_lib.prof_enable("gemm")
rt._conv3_wgrad(dy, x16, c, c, dims, 1)
ms, work, n = _lib.prof_collect("gemm")
_lib.prof_enable()
print("   GEMM alone ms", ms, "launches", n)
T = 32768
a_t = torch.randn(512, T, device=DEV).to(BF); b_t = torch.randn(2048, T, device=DEV).to(BF)
out = torch.zeros(512, 2048, device=DEV)
s = _lib.cur_stream()
print("dW_ff2 [512 x 2048, K = 32768] ms", timed(lambda: _lib.call("rald_gemm_bf16_accum", a_t.data_ptr(), T, b_t.data_ptr(), T, out.data_ptr(), 2048, 512, 2048, T, s), 10))
a_t = torch.randn(512, T, device=DEV).to(BF); b_t = torch.randn(512, T, device=DEV).to(BF)
out = torch.zeros(512, 512, device=DEV)
print("dW_o [512 x 512, K = 32768] ms", timed(lambda: _lib.call("rald_gemm_bf16_accum", a_t.data_ptr(), T, b_t.data_ptr(), T, out.data_ptr(), 512, 512, 512, T, s), 10))
