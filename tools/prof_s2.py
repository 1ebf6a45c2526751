"""Stride-2 exchange conv at the bench shape (batch 16, 32 ch, 16x64x160 -> 8x32x80) through the space-to-depth view:
forward, dgrad, wgrad — for ncu captures and CUDA-event timing."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import ops  # noqa: E402
from rtpose_b200.p8 import P8  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N, C, grid = 16, 32, (16, 64, 160)
og = (8, 32, 80)
torch.manual_seed(0)
x = P8.from_ncdhw(torch.randn(N, C, *grid, device="cuda"))
dy = P8.from_ncdhw(torch.randn(N, C, *og, device="cuda"))
w = torch.randn(C, C, 3, 3, 3, device="cuda") * 0.03
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
packs = ops.PackedWeights()
st = ops.gn_stats(x, 8)
xs = ops.gn_apply_s2d(x, 8, st, gamma, beta, P8(N, 8 * C, *og))
we = ops.s2d_expand(w)
y = P8(N, C, *og)
dxs = P8(N, 8 * C, *og)
gw = torch.zeros_like(w)
masks = [ops.s2d_tap_mask(p, False) for p in range(8)]


def timeit(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


timeit("gn_apply_s2d", lambda: ops.gn_apply_s2d(x, 8, st, gamma, beta, xs))
timeit("s2 fwd  (conv_k3s1 over the view, masked)", lambda: ops.conv_forward(packs, xs, we, 1, y, key="k", version=0, tap_mask=masks))
timeit("s2 dgrad (4 paired conv_k3s1 launches)", lambda: ops.conv_dgrad(packs, dy, we, 1, dxs, key="k", version=0, s2d_cin=C))
timeit("s2 wgrad (gather kernel over the view)", lambda: ops.conv_wgrad_s2d(xs, dy, C, gw))
