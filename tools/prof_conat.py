"""Runs rtp_conat_fwd at the bench shape (batch 16, 16x64x160, 32+32+64+64 -> 128) a few times: the target of
`ncu --set full -k regex:conat_kernel` (profiles/r02_ncu_conat.txt) and a CUDA-event timing."""
import sys

import torch

sys.path.insert(0, ".")
from rtpose_b200 import ops  # noqa: E402
from rtpose_b200.p8 import P8  # noqa: E402

N, grid, chans, Cout = 16, (16, 64, 160), (32, 32, 64, 64), 128
g = torch.Generator(device="cuda").manual_seed(1)
ys = []
for j, c in enumerate(chans):
    gr = tuple(v >> j for v in grid)
    ys.append(P8.from_ncdhw(torch.randn(N, c, *gr, device="cuda", generator=g)))
w = torch.randn(Cout, sum(chans), 1, 1, 1, device="cuda", generator=g) * 0.1
b = torch.randn(Cout, device="cuda", generator=g)
packs = ops.PackedWeights()
out = P8(N, Cout, *grid)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    assert ops.conat_forward(packs, ys, w, b, out) is not None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conat_forward(packs, ys, w, b, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = 16.0 * N * (chans[0] // 8 + Cout // 8) * grid[0] * grid[1] * grid[2]
print("conat_fwd: %.4f ms per launch, %.1f GB/s algorithmic (x0 read + out written)" % (ms, nbytes / ms / 1e6))
