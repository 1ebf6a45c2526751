"""Stage split of the tensor-core DCN forward (sampler / contraction / unpack)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtpose_b200.dcn as D
from rtpose_b200 import ops
D.TENSOR_CORE = True
N, C, H, W, dg = 16, int(os.environ.get("C", "128")), 64, 160, 4
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, C, H, W, device="cuda", generator=g)
w = torch.randn(C, C, 3, 3, device="cuda", generator=g) * 0.05
for name, scale in (("random offsets sigma 1.5", 1.5), ("smooth offsets (zero)", 0.0)):
    off = torch.randn(N, dg * 18, H, W, device="cuda", generator=g) * scale
    D.deform_conv(x, off, w, 1, 1, 1, 1, dg); torch.cuda.synchronize()
    ops.PROFILE = {}
    import rtpose_b200.lib as L
    evs = []
    orig = L.call
    def timed(name_, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name_, *a); e1.record(); evs.append((name_, e0, e1))
    L.call = timed
    for _ in range(3):
        D.deform_conv(x, off, w, 1, 1, 1, 1, dg)
    torch.cuda.synchronize(); L.call = orig; ops.PROFILE = None
    agg = {}
    for n_, e0, e1 in evs:
        agg[n_] = agg.get(n_, 0.0) + e0.elapsed_time(e1) / 3
    print(name, {k: round(v, 3) for k, v in agg.items()})
