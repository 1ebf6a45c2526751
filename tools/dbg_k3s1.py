import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtpose_b200 import ops, lib
from rtpose_b200.p8 import P8, _stream
N, Cc, grid = 16, 32, (16, 64, 160)
x = P8.from_ncdhw(torch.randn(N, Cc, *grid, device="cuda"))
w = torch.randn(Cc, Cc, 3, 3, 3, device="cuda") * 0.03
out = P8(N, Cc, *grid)
packs = ops.PackedWeights()
wp = packs.get_k3s1(w, 32, 32, False)
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
d = lib.ConvK3S1Desc()
d.inp, d.out, d.res, d.mask = x.struct(), out.struct(), lib.NULL_P8, lib.NULL_P8
d.w, d.bias, d.Cin, d.NPo, d.out_c8, d.relu, d.accumulate = wp.data_ptr(), None, 32, 32, 4, 0, 0
for i in range(3):
    d.debug = dbg.data_ptr()
    lib.call("rtp_conv_k3s1", C.byref(d), _stream())
torch.cuda.synchronize()
t = dbg.view(148, 8).float()
m = t.mean(0)
print("per CTA: total %.0f cyc, wait acc_empty %.0f, wait full %.0f, issue %.0f, plane-steps %.0f" % tuple(m[:5].tolist()))
print("per plane-step: total %.0f, empty %.0f, full %.0f, tap-loop issue %.0f, first-touch %.0f, commits+syncwarp %.0f" % tuple((m[[0, 1, 2, 3, 5, 6]] / m[4]).tolist()))
print("(needs a build with RTP_NVCC_EXTRA=-DRTP_K3S1_DEBUG python -m rtpose_b200.build --force)")
