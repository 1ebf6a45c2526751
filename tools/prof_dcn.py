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

# ---- the reference's own extension (det3d/ops/dcn/src), built UNMODIFIED from the staged copy baseline/_ref with
#      -DAT_CHECK=TORCH_CHECK (the macro was renamed in torch >= 1.5); outputs only under baseline/_ref/_build.  "The kernel to
#      beat" of SURVEY §2.2.  Skipped (with the reason) when the staged sources are absent or do not compile on this torch.
if "--ref-ext" in sys.argv:
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "det3d", "ops", "dcn", "src")
    try:
        from torch.utils.cpp_extension import load
        bdir = os.path.join(os.path.dirname(src), "_build")
        os.makedirs(bdir, exist_ok=True)
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        ext = load(name="deform_conv_cuda_ref", sources=[os.path.join(src, "deform_conv_cuda.cpp"), os.path.join(src, "deform_conv_cuda_kernel.cu")],
                   extra_cflags=["-DAT_CHECK=TORCH_CHECK", "-w"], extra_cuda_cflags=["-DAT_CHECK=TORCH_CHECK", "-w"], build_directory=bdir, verbose=False)
    except Exception as ex:  # noqa: BLE001
        print("reference DCN extension: not built (%s)" % (repr(ex)[:400],))
        ext = None
    if ext is not None:
        for (N, C) in ((16, 128), (16, 256)):
            H, W, dg, step = 64, 160, 4, 16
            g = torch.Generator(device="cuda").manual_seed(0)
            x = torch.randn(N, C, H, W, device="cuda", generator=g)
            off = torch.randn(N, dg * 18, H, W, device="cuda", generator=g)
            w = torch.randn(C, C, 3, 3, device="cuda", generator=g) * 0.05
            gy = torch.randn(N, C, H, W, device="cuda", generator=g)
            out = x.new_empty(N, C, H, W)
            bufs = [x.new_empty(0), x.new_empty(0)]

            def fwd():  # deform_conv.py:48-67
                ext.deform_conv_forward_cuda(x, w, off, out, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, step)

            def bwd(pstep):  # deform_conv.py:84-110
                gi, go, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
                ext.deform_conv_backward_input_cuda(x, off, gy, gi, go, w, bufs[0], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, step)
                ext.deform_conv_backward_parameters_cuda(x, off, gy, gw, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, 1, pstep)
            t_f = bench(fwd)
            ref = tv.deform_conv2d(x, off, w, None, padding=1)
            err = float((out - ref).abs().max() / ref.abs().max())
            print("N=%d C=%d reference extension (im2col_step %d) fwd %.3f ms  %.1f TFLOP/s (max rel diff to torchvision %.2e)" %
                  (N, C, step, t_f, 2.0 * N * H * W * C * C * 9 / 1e9 / t_f, err))
            for pstep in (step, 1):
                # with the wrapper's im2col_step (min(64, N)) the parameter gradient fails on torch >= 1.5: zeros_like() of the
                # transposed gradOutput keeps its strides and the following .view() throws (deform_conv_cuda.cpp:395-411); only
                # im2col_step = 1 runs.  The sources are not modified.
                try:
                    t_fb = bench(lambda: (fwd(), bwd(pstep)), 3)
                    print("N=%d C=%d reference extension fwd+bwd %.3f ms (backward_parameters im2col_step %d)" % (N, C, t_fb, pstep))
                    break
                except RuntimeError as ex:
                    print("N=%d C=%d reference extension backward_parameters with im2col_step %d fails on this torch: %s" %
                          (N, C, pstep, str(ex)[:120]))
