"""DCN timing at FeatureAdaption scale (center_head.py:24-62: DeformConv(C, C, 3, pad 1, dg 4) on [B*16, C, 64, 160])."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtpose_b200.dcn import deform_conv
import torchvision.ops as tv

def bench(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (N, C) in ((16, 128), (32, 128), (16, 256)):
    H, W, dg = 64, 160, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    off = torch.randn(N, dg * 18, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) * 0.05
    gf = 2.0 * N * H * W * C * C * 9 / 1e9
    import rtpose_b200.dcn as D
    for name, tc in (("fp32 CUDA cores", False), ("tcgen05 bf16", True)):
        D.TENSOR_CORE = tc
        t = bench(lambda: deform_conv(x, off, w, 1, 1, 1, 1, dg))
        print("N=%d C=%d ours[%s] fwd %.3f ms  %.1f TFLOP/s" % (N, C, name, t, gf / t))
    D.TENSOR_CORE = False
    t = bench(lambda: tv.deform_conv2d(x, off, w, None, padding=1))
    print("N=%d C=%d torchvision fwd %.3f ms  %.1f TFLOP/s" % (N, C, t, gf / t))
    xr, orr, wr = x.clone().requires_grad_(True), off.clone().requires_grad_(True), w.clone().requires_grad_(True)
    def fb(fn):
        for t_ in (xr, orr, wr): t_.grad = None
        fn().sum().backward()
    for name, tc in (("fp32 CUDA cores", False), ("tcgen05 bf16", True)):
        D.TENSOR_CORE = D.TENSOR_CORE_BACKWARD = tc
        t = bench(lambda: fb(lambda: deform_conv(xr, orr, wr, 1, 1, 1, 1, dg)), 3)
        print("N=%d C=%d ours[%s] fwd+bwd %.3f ms" % (N, C, name, t))
    D.TENSOR_CORE = D.TENSOR_CORE_BACKWARD = False
    t = bench(lambda: fb(lambda: tv.deform_conv2d(xr, orr, wr, None, padding=1)), 3)
    print("N=%d C=%d torchvision fwd+bwd %.3f ms" % (N, C, t))
