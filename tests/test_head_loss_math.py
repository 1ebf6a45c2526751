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


def test_targets_on_one_voxel_add_their_regression_gradients():
    """Two targets of a sample may land on the same voxel (different joints / classes).  _transpose_and_gather_feat gathers that
    voxel twice, so autograd of the reference loss ADDS both contributions there, and a masked-out target adds nothing.  This is
    what head_loss_final_kernel's per-voxel accumulation chain computes (csrc/head.cu: for target i, the bf16 running sum over
    the sample's targets j with ind[j] == ind[i], in order — the value a read-modify-write per target leaves behind)."""
    grid, N, ncls, R, M = (4, 6, 8), 2, 15, 3, 15
    rs = np.random.RandomState(3)
    V = grid[0] * grid[1] * grid[2]
    hm = torch.from_numpy(rs.randn(N, ncls, *grid)).requires_grad_(True)
    reg = torch.from_numpy(rs.randn(N, R, *grid)).requires_grad_(True)
    ind = np.stack([rs.choice(V, size=M, replace=False) for _ in range(N)])
    ind[0, 4] = ind[0, 3]            # joints 3 and 4 of sample 0 share a voxel
    ind[1, 7] = ind[1, 2]            # joints 2 and 7 of sample 1 too, and joint 7 is masked out
    mask = np.ones((N, M), np.uint8)
    mask[1, 7] = 0
    tgt = {"hm": torch.from_numpy(rs.rand(N, ncls, *grid) ** 4), "ind": torch.from_numpy(ind), "mask": torch.from_numpy(mask),
           "cat": torch.from_numpy(np.tile(np.arange(M), (N, 1))), "anno_pose": torch.from_numpy(rs.randn(N, M, R))}
    weight, code_w = 0.5, [1.0, 1.5, 2.0]
    O.head_loss({"hm": hm, "reg": reg}, tgt, weight, code_w)["loss"].backward()
    den = float(mask.sum()) + 1e-4
    want = torch.zeros_like(reg)
    for n in range(N):
        for m in range(M):
            if not mask[n, m]:
                continue
            idx = int(ind[n, m])
            z, y, x = idx // (grid[1] * grid[2]), (idx // grid[2]) % grid[1], idx % grid[2]
            want[n, :, z, y, x] += weight * torch.tensor(code_w) * torch.sign(reg.detach()[n, :, z, y, x] - tgt["anno_pose"][n, m]) / den
    torch.testing.assert_close(want, reg.grad, rtol=1e-6, atol=1e-12)
    # the shared voxel of sample 0 carries two contributions, that of sample 1 one (the masked joint adds nothing)
    i0, i1 = int(ind[0, 3]), int(ind[1, 2])
    g0 = reg.grad[0].reshape(R, -1)[:, i0].abs() * den / (weight * torch.tensor(code_w))
    g1 = reg.grad[1].reshape(R, -1)[:, i1].abs() * den / (weight * torch.tensor(code_w))
    assert set(np.round(g0.numpy()).astype(int).tolist()) <= {0, 2} and np.allclose(g1.numpy(), 1.0)
