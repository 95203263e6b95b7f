"""Dev check on a B200: tcgen05 attention, LN rows vs torch fp32."""
import sys
sys.path.insert(0, ".")
import torch
from rald_b200 import _lib
from rald_b200._lib import c_void_p, c_int, c_i64, c_f32
L = _lib.lib()
L.rald_attn_d64.argtypes = [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_f32, c_void_p]
L.rald_ln_rows.argtypes = [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_i64, c_int, c_i64, c_int, c_f32, c_void_p]
dev = "cuda"
torch.manual_seed(0)

def attn(B, H, Sq, Skv, amp=1.0, timing=False, ramp=0.0):
    D = H * 64
    q = (torch.randn(B * Sq, D, device=dev) * amp).bfloat16()
    k = torch.randn(B * Skv, D, device=dev) * amp
    if ramp:   # key magnitude grows with the key index: later key chunks raise the running maximum (lazy-shift path)
        k = (k.view(B, Skv, D) * torch.linspace(0.05, ramp, Skv, device=dev)[None, :, None]).reshape(B * Skv, D)
    k = k.bfloat16()
    v = torch.randn(B * Skv, D, device=dev).half()   # V is fp16 (see include/rald_b200.h)
    o = torch.zeros(B * Sq, D, device=dev, dtype=torch.bfloat16)
    st = _lib.cur_stream()
    args = (q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, B, H, Sq, Skv, 0.125, st)
    _lib.check(L.rald_attn_d64(*args), "attn")
    torch.cuda.synchronize()
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2)
    kf = k.float().view(B, Skv, H, 64).transpose(1, 2)
    vf = v.float().view(B, Skv, H, 64).transpose(1, 2)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, -1) @ vf
    ref = ref.transpose(1, 2).reshape(B * Sq, D)
    err = ((o.float() - ref).norm() / ref.norm()).item()
    print(f"attn B={B} H={H} Sq={Sq} Skv={Skv} amp={amp} ramp={ramp}: rel={err:.3e} {'OK' if err < 1e-2 else 'FAIL'}", flush=True)
    if timing:
        for _ in range(3): L.rald_attn_d64(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): L.rald_attn_d64(*args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 4.0 * B * H * Sq * Skv * 64
        print(f"   {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
    return err < 1e-2

def ln(rows, per_frame):
    x = torch.randn(rows, 512, device=dev) * 3 + 1
    nf = rows // per_frame if per_frame else 1
    mod = torch.randn(nf, 1024, device=dev)
    out = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
    _lib.check(L.rald_ln_rows(x.data_ptr(), 512, mod.data_ptr(), mod.data_ptr() + 512 * 4, 1024 if per_frame else 0,
                              per_frame, 1, out.data_ptr(), 512, 0, rows, 512, 1e-5, _lib.cur_stream()), "ln")
    torch.cuda.synchronize()
    xn = torch.nn.functional.layer_norm(x, (512,))
    if per_frame:
        ref = xn.view(nf, per_frame, 512) * (1 + mod[:, None, :512]) + mod[:, None, 512:]
        ref = ref.reshape(rows, 512)
    else:
        ref = xn * (1 + mod[0, :512]) + mod[0, 512:]
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"ln rows={rows} per_frame={per_frame}: rel={err:.3e} {'OK' if err < 4e-3 else 'FAIL'}", flush=True)
    return err < 4e-3

ok = True
ok &= attn(1, 1, 128, 64)
ok &= attn(2, 3, 256, 384)
ok &= attn(5, 8, 512, 256)
ok &= attn(1, 1, 128, 128)
ok &= attn(1, 1, 128, 256)
ok &= attn(1, 1, 128, 512)
ok &= attn(1, 8, 512, 512)
ok &= attn(2, 8, 512, 64)
ok &= attn(3, 8, 512, 512, amp=3.0)
ok &= attn(2, 8, 512, 512, amp=5.0)
ok &= attn(2, 8, 512, 64, amp=4.0)
ok &= attn(2, 8, 512, 512, amp=2.0, ramp=4.0)
ok &= attn(1, 8, 512, 512, amp=2.0, ramp=4.0)
ok &= attn(3, 2, 256, 512, amp=1.0, ramp=8.0)
ok &= attn(64, 8, 512, 512, timing=True)
ok &= attn(64, 8, 512, 64, timing=True)
ok &= attn(8, 8, 512, 512, timing=True)
ok &= ln(512, 0)
ok &= ln(2048, 512)
ok &= ln(32768, 512)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
