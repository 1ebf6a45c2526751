"""Runs the dominant kernels in isolation at the bench shape (batch 16, 32 ch, 16x64x160) — for ncu captures and
CUDA-event timing:  python tools/prof_kernels.py [conv|wgrad|gn|all] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import ops  # noqa: E402
from rtpose_b200.p8 import P8  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N, C, grid = 16, 32, (16, 64, 160)
torch.manual_seed(0)
x = P8.from_ncdhw(torch.randn(N, C, *grid, device="cuda"))
dy = P8.from_ncdhw(torch.randn(N, C, *grid, device="cuda"))
w = torch.randn(C, C, 3, 3, 3, device="cuda") * 0.03
out = P8(N, C, *grid)
packs = ops.PackedWeights()
dW = torch.zeros_like(w)
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")


def timeit(name, fn, flops=None, bytes_=None):
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
    msg = "%-28s %8.3f ms" % (name, ms)
    if flops:
        msg += "  %7.1f TFLOP/s" % (flops / ms / 1e9)
    if bytes_:
        msg += "  %7.1f GB/s" % (bytes_ / ms / 1e6)
    print(msg)


fl = 2.0 * N * grid[0] * grid[1] * grid[2] * C * C * 27
tb = N * 4 * grid[0] * grid[1] * grid[2] * 16  # bytes of one 32-channel tensor (real voxels)
if what in ("conv", "all"):
    timeit("conv_k3s1 fwd 32->32", lambda: ops.conv_forward(packs, x, w, 1, out), fl)
    timeit("conv_k3s1 fwd +res+relu", lambda: ops.conv_forward(packs, x, w, 1, out, relu=True, res=dy), fl)
    timeit("conv_k3s1 dgrad", lambda: ops.conv_dgrad(packs, dy, w, 1, out), fl)
    stx = ops.gn_stats(x, 8)
    timeit("conv_k3s1 fwd +res+relu +stats", lambda: ops.conv_forward(packs, x, w, 1, out, relu=True, res=dy, stat=("stats", 8, 1e-5)), fl)
    timeit("conv_k3s1 dgrad +gn-bwd sums", lambda: ops.conv_dgrad(packs, dy, w, 1, out, stat=("red", 8, x, stx)), fl)
    ops.USE_K3S1 = False
    timeit("conv_generic fwd 32->32", lambda: ops.conv_forward(packs, x, w, 1, out), fl)
    ops.USE_K3S1 = True
if what in ("wgrad", "all"):
    timeit("wgrad_k3s1 32->32", lambda: ops.conv_wgrad(x, dy, 3, 1, dW), fl)
if what in ("gn", "all"):
    st = ops.gn_stats(x, 8)
    timeit("gn_stats", lambda: ops.gn_stats(x, 8), bytes_=tb)
    timeit("gn_apply", lambda: ops.gn_apply(x, 8, st, gamma, beta, out), bytes_=2 * tb)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    timeit("gn_backward (reduce+apply)", lambda: ops.gn_backward(x, dy, 8, st, gamma, dg, db, False, out, False), bytes_=5 * tb)
    timeit("grad_add", lambda: ops.grad_add(dy, out, mask=x, accumulate=True), bytes_=4 * tb)
