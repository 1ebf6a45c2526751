"""Host-side check of the closed-form gradients rtp_head_loss writes together with the loss (csrc/head.cu) against autograd
of the oracle's restatement of CenterHead.loss (center_head.py:244-270, FastFocalLoss / RegLoss centernet_loss.py:17-54), in
float64: with p = clamp(sigmoid(h), 1e-4, 1 - 1e-4),
  d(-neg/np)/dh = -(1/np) (2 p log(1-p) - p^2/(1-p)) (1-t)^4 p (1-p)          at every voxel (0 where the clamp is active),
  d(-pos/np)/dh = -(1/np) ((1-p)^2/p - 2 (1-p) log p) p (1-p)                 at the positive locations,
  d loss / d reg = weight * code_w * sign(reg - anno) / (sum(mask) + 1e-4)     at the positive locations."""
import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O


@pytest.mark.parametrize("one_hm", [True, False])
def test_loss_gradient_closed_forms(one_hm):
    grid, N = (4, 6, 8), 3
    ncls, R, M = (1, 45, 1) if one_hm else (15, 3, 15)
    rs = np.random.RandomState(ncls)
    V = grid[0] * grid[1] * grid[2]
    hm = torch.from_numpy(rs.randn(N, ncls, *grid) * 3).requires_grad_(True)
    hm.data[0, 0, 0, 0, 0], hm.data[1, 0, 1, 1, 1] = 12.0, -12.0   # both clamps active somewhere
    reg = torch.from_numpy(rs.randn(N, R, *grid)).requires_grad_(True)
    tgt = {"hm": torch.from_numpy(rs.rand(N, ncls, *grid) ** 4), "ind": torch.from_numpy(rs.randint(0, V, size=(N, M))),
           "mask": torch.from_numpy((rs.rand(N, M) > 0.3).astype(np.uint8)), "cat": torch.from_numpy(np.tile(np.arange(M) % ncls, (N, 1))),
           "anno_pose": torch.from_numpy(rs.randn(N, M, R))}
    for n in range(N):  # unique voxels per sample so that no two positives share an element
        tgt["ind"][n] = torch.from_numpy(rs.choice(V, size=M, replace=False))
    weight, code_w = 0.5, list(np.linspace(1.0, 2.0, R))
    O.head_loss({"hm": hm, "reg": reg}, tgt, weight, code_w)["loss"].backward()

    h = hm.detach()
    s = torch.sigmoid(h)
    p = s.clamp(1e-4, 1 - 1e-4)
    inside = ((s >= 1e-4) & (s <= 1 - 1e-4)).double()
    npos = float(tgt["mask"].sum())
    assert npos > 0
    g = -(1 / npos) * (2 * p * torch.log(1 - p) - p * p / (1 - p)) * (1 - tgt["hm"]) ** 4 * p * (1 - p) * inside
    g_reg = torch.zeros_like(reg)
    den = npos + 1e-4
    for n in range(N):
        for m in range(M):
            if not tgt["mask"][n, m]:
                continue
            idx = int(tgt["ind"][n, m])
            z, y, x = idx // (grid[1] * grid[2]), (idx // grid[2]) % grid[1], idx % grid[2]
            c = int(tgt["cat"][n, m])
            pp = p[n, c, z, y, x]
            g[n, c, z, y, x] += -(1 / npos) * ((1 - pp) ** 2 / pp - 2 * (1 - pp) * torch.log(pp)) * pp * (1 - pp) * inside[n, c, z, y, x]
            g_reg[n, :, z, y, x] = weight * torch.tensor(code_w) * torch.sign(reg.detach()[n, :, z, y, x] - tgt["anno_pose"][n, m]) / den
    torch.testing.assert_close(g, hm.grad, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(g_reg, reg.grad, rtol=1e-6, atol=1e-12)  # the reference forms sum(mask) + 1e-4 in float32
    assert float(hm.grad[0, 0, 0, 0, 0]) == 0.0 and float(hm.grad[1, 0, 1, 1, 1]) == 0.0  # clamped logits get no gradient
