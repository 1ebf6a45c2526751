"""Branch-exchange kernels at the bench shape: upsample backward (3 separable passes) and fuse_sum."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import ops  # noqa: E402
from rtpose_b200.p8 import P8  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N = 16


def timeit(name, fn, mb=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-52s %8.3f ms%s" % (name, ms, "  %7.1f GB/s" % (mb / ms) if mb else ""))


for C in (32, 128):
    g = P8.from_ncdhw(torch.randn(N, C, 16, 64, 160, device="cuda"))
    mb_full = N * C * 16 * 64 * 160 * 2 / 1e6
    for f in (2, 4, 8):
        low = P8(N, C, 16 // f, 64 // f, 160 // f)
        timeit("upsample_bwd %3d ch full -> 1/%d" % (C, f), lambda: ops.upsample_bwd(g, low), mb_full * (1 + 1.0 / f))
    same = P8.from_ncdhw(torch.randn(N, C, 16, 64, 160, device="cuda"))
    lows = [P8.from_ncdhw(torch.randn(N, C, 16 // f, 64 // f, 160 // f, device="cuda")) for f in (2, 4, 8)]
    out = P8(N, C, 16, 64, 160)
    timeit("fuse_sum %3d ch: same + 3 low + relu" % C, lambda: ops.fuse_sum(out, [same], lows, relu=True), 2 * mb_full)
    timeit("fuse_sum %3d ch: same + 2 low + relu" % C, lambda: ops.fuse_sum(out, [same], lows[:2], relu=True), 2 * mb_full)
    timeit("grad_add %3d ch (mask, accumulate)" % C, lambda: ops.grad_add(g, out, mask=same, accumulate=True), 4 * mb_full)
