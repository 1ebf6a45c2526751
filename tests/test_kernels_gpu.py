"""GPU parity tests, kernel level: every C-ABI kernel against the PyTorch (CPU, fp32) op the reference calls at
that site, on identical bf16-rounded inputs.  Tolerances (SURVEY.md §8c contract (1)):
  convs / wgrad : |d| <= 2^-7 * max|ref|   (one bf16 ulp of the output range; fp32 accumulation on both sides)
  fp32 formulas : rel 1e-4 (GroupNorm statistics, loss)
  indices       : exact
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_ULP = 2.0 ** -7


def bf(x):
    return x.to(torch.bfloat16).float()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return bf(torch.randn(*shape, generator=g) * scale)


def to_p8(x):
    from rtpose_b200.p8 import P8
    return P8.from_ncdhw(x.cuda())


def close(got, ref, tol=BF16_ULP, what=""):
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    err = (got - ref).abs().max().item()
    lim = tol * max(ref.abs().max().item(), 1e-6)
    assert err <= lim, "%s: max abs err %.4g > %.4g (ref max %.4g)" % (what, err, lim, ref.abs().max().item())


@pytest.fixture(scope="module")
def ctx():
    from rtpose_b200 import lib, ops
    lib.require_device()
    return ops.PackedWeights()


def test_pack_unpack_roundtrip():
    x = rnd(2, 13, 4, 10, 12, seed=1)
    t = to_p8(x)
    y = t.to_ncdhw().cpu()
    assert torch.equal(y, x)
    # pads stay zero
    full = t.buf[t.offset:t.offset + t.N * t.n_stride].view(t.N, t.C8, t.Z, t.X + 2, t.Y + 2, 8).float().cpu()
    assert full[:, :, :, 0].abs().max() == 0 and full[:, :, :, -1].abs().max() == 0
    assert full[:, :, :, :, 0].abs().max() == 0 and full[:, :, :, :, -1].abs().max() == 0
    assert full[:, 1, :, :, :, 5:].abs().max() == 0  # channels 13..15


CONV_CASES = [
    # N, Cin, Cout, grid, k, stride, bias, relu, res
    (2, 32, 32, (4, 10, 12), 3, 1, False, True, False),
    (1, 32, 32, (5, 9, 7), 3, 1, False, True, True),
    (2, 32, 64, (4, 10, 12), 3, 2, False, True, False),
    (1, 64, 64, (3, 6, 10), 3, 1, False, False, False),
    (2, 64, 32, (2, 4, 6), 1, 1, False, False, False),
    (1, 128, 64, (3, 8, 6), 3, 1, True, True, False),
    (2, 32, 45, (4, 6, 8), 3, 1, True, False, False),
    (2, 32, 1, (4, 6, 8), 3, 1, True, False, False),
    (1, 32, 32, (5, 7, 9), 3, 2, False, False, False),
    (1, 192, 128, (2, 4, 6), 1, 1, True, False, False),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv_forward(ctx, case):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid, k, stride, use_bias, relu, use_res = case
    x = rnd(N, Cin, *grid, seed=2)
    w = rnd(Cout, Cin, k, k, k, seed=3, scale=(Cin * k ** 3) ** -0.5)
    b = rnd(Cout, seed=4) if use_bias else None
    ref = F.conv3d(x, w, b, stride=stride, padding=k // 2)
    res = rnd(*ref.shape, seed=5) if use_res else None
    if res is not None:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    xp = to_p8(x)
    out = P8(N, Cout, *ref.shape[2:])
    wc = w.cuda()
    ops.conv_forward(ctx, xp, wc, stride, out, bias=b.cuda() if use_bias else None, relu=relu,
                     res=to_p8(res) if use_res else None)
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, what="conv fwd")


DGRAD_CASES = [(2, 32, 32, (4, 10, 12), 3, 1), (1, 32, 64, (4, 10, 12), 3, 2), (1, 64, 64, (5, 7, 9), 3, 2),
               (2, 64, 32, (2, 4, 6), 1, 1), (1, 128, 64, (3, 8, 6), 3, 1), (1, 32, 45, (3, 6, 8), 3, 1)]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[str(c) for c in DGRAD_CASES])
def test_conv_dgrad_and_wgrad(ctx, case):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid, k, stride = case
    x = rnd(N, Cin, *grid, seed=6).requires_grad_(True)
    w = rnd(Cout, Cin, k, k, k, seed=7, scale=(Cin * k ** 3) ** -0.5).requires_grad_(True)
    y = F.conv3d(x, w, None, stride=stride, padding=k // 2)
    dy = rnd(*y.shape, seed=8)
    y.backward(dy)
    # dgrad, with ReLU mask and accumulate variants
    dyp = to_p8(dy)
    dx = P8(N, Cin, *grid)
    ops.conv_dgrad(ctx, dyp, w.detach().cuda(), stride, dx)
    torch.cuda.synchronize()
    close(dx.to_ncdhw(), x.grad, what="dgrad")
    mask = rnd(N, Cin, *grid, seed=9)
    base = rnd(N, Cin, *grid, seed=10)
    dx2 = to_p8(base)
    ops.conv_dgrad(ctx, dyp, w.detach().cuda(), stride, dx2, mask=to_p8(mask), accumulate=True)
    torch.cuda.synchronize()
    close(dx2.to_ncdhw(), base + x.grad * (mask > 0), tol=2 * BF16_ULP, what="dgrad mask+acc")
    # wgrad
    dW = torch.full(w.shape, 7.0, device="cuda")
    ops.conv_wgrad(to_p8(x.detach()), dyp, k, stride, dW)
    torch.cuda.synchronize()
    close(dW, w.grad, tol=2e-3, what="wgrad")
    ops.conv_wgrad(to_p8(x.detach()), dyp, k, stride, dW, accumulate=True)
    torch.cuda.synchronize()
    close(dW, 2 * w.grad, tol=2e-3, what="wgrad accumulate")


def test_wgrad_channel_slices(ctx):
    """final-conv style input-channel slices and merged-head style output-channel offsets"""
    from rtpose_b200 import ops
    x = rnd(2, 32, 3, 6, 8, seed=11)
    dy = rnd(2, 64, 3, 6, 8, seed=12)
    w = torch.zeros(64, 32, 3, 3, 3, requires_grad=True)
    F.conv3d(x, w, padding=1).backward(dy)
    g_lo = torch.zeros(32, 32, 3, 3, 3, device="cuda")
    g_hi = torch.zeros(32, 32, 3, 3, 3, device="cuda")
    ops.conv_wgrad(to_p8(x), to_p8(dy), 3, 1, g_lo, n0=0, more=((g_hi, False, 0, 32),))
    torch.cuda.synchronize()
    close(g_lo, w.grad[:32], tol=2e-3, what="wgrad n0=0")
    close(g_hi, w.grad[32:], tol=2e-3, what="wgrad n0=32")
    big = torch.zeros(64, 96, 3, 3, 3, device="cuda")
    ops.conv_wgrad(to_p8(x), to_p8(dy), 3, 1, big, ci0=64)
    torch.cuda.synchronize()
    close(big[:, 64:], w.grad, tol=2e-3, what="wgrad ci0")
    assert big[:, :64].abs().max().item() == 0


@pytest.mark.parametrize("C,grid", [(32, (4, 10, 12)), (64, (2, 5, 7)), (128, (2, 4, 6))])
def test_groupnorm_fwd_bwd(C, grid):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, G = 2, 8
    x = (rnd(N, C, *grid, seed=13) + 0.5).requires_grad_(True)
    x.data = bf(x.data)
    gamma = (1 + 0.2 * torch.randn(C)).requires_grad_(True)
    beta = (0.1 * torch.randn(C)).requires_grad_(True)
    y = F.group_norm(x, G, gamma, beta, 1e-5)
    dy = rnd(*y.shape, seed=14)
    y.backward(dy)
    xp = to_p8(x.detach())
    stats = ops.gn_stats(xp, G)
    xr = x.detach().reshape(N, G, -1)
    torch.cuda.synchronize()
    np.testing.assert_allclose(stats[..., 0].cpu().numpy(), xr.mean(-1).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stats[..., 1].cpu().numpy(), (xr.var(-1, unbiased=False) + 1e-5).rsqrt().numpy(), rtol=1e-4)
    out = ops.gn_apply(xp, G, stats, gamma.detach().cuda(), beta.detach().cuda(), P8(N, C, *grid))
    torch.cuda.synchronize()
    close(out.to_ncdhw(), y, what="gn apply")
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx = P8(N, C, *grid)
    ops.gn_backward(xp, to_p8(dy), G, stats, gamma.detach().cuda(), dg, db, False, dx, False)
    torch.cuda.synchronize()
    close(dx.to_ncdhw(), x.grad, what="gn dx")
    close(dg, gamma.grad, tol=1e-3, what="gn dgamma")
    close(db, beta.grad, tol=1e-3, what="gn dbeta")
    # relu-masked accumulate variant
    xp.relu_out = True
    base = rnd(N, C, *grid, seed=15)
    dx2 = to_p8(base)
    ops.gn_backward(xp, to_p8(dy), G, stats, gamma.detach().cuda(), dg, db, True, dx2, True)
    torch.cuda.synchronize()
    close(dx2.to_ncdhw(), base + x.grad * (x.detach() > 0), tol=2 * BF16_ULP, what="gn dx mask+acc")
    close(dg, 2 * gamma.grad, tol=1e-3, what="gn dgamma acc")
    # a second gradient into x folded into the same pass (`add`): dx = [x > 0] * (gn term + add) (+ old)
    extra = rnd(N, C, *grid, seed=16)
    dx3, dx4 = P8(N, C, *grid), to_p8(base)
    ops.gn_backward(xp, to_p8(dy), G, stats, gamma.detach().cuda(), dg, db, False, dx3, False, add=to_p8(extra))
    ops.gn_backward(xp, to_p8(dy), G, stats, gamma.detach().cuda(), dg, db, False, dx4, True, add=to_p8(extra))
    torch.cuda.synchronize()
    want = (x.grad + bf(extra)) * (x.detach() > 0)
    close(dx3.to_ncdhw(), want, tol=2 * BF16_ULP, what="gn dx + add, masked")
    close(dx4.to_ncdhw(), base + want, tol=3 * BF16_ULP, what="gn dx + add, masked, accumulated")


def test_fuse_sum_and_upsample_bwd():
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, C = 2, 32
    hi, l1, l2 = (8, 16, 24), (4, 8, 12), (1, 2, 3)
    a, b2 = rnd(N, C, *hi, seed=16), rnd(N, C, *hi, seed=17)
    u1 = rnd(N, C, *l1, seed=18).requires_grad_(True)
    u2 = rnd(N, C, *l2, seed=19).requires_grad_(True)
    bias = rnd(C, seed=20)
    pre = a + b2 + F.interpolate(u1, size=hi, mode="trilinear", align_corners=True) + \
        F.interpolate(u2, size=hi, mode="trilinear", align_corners=True) + bias.view(1, C, 1, 1, 1)
    ref = F.relu(pre)
    out = ops.fuse_sum(P8(N, C, *hi), [to_p8(a), to_p8(b2)], [to_p8(u1.detach()), to_p8(u2.detach())],
                       bias=bias.cuda(), relu=True)
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, what="fuse_sum")
    g = rnd(N, C, *hi, seed=21)
    pre.backward(g)
    gp = to_p8(g)
    for u, shape in ((u1, l1), (u2, l2)):
        d = P8(N, C, *shape)
        ops.upsample_bwd(gp, d)
        torch.cuda.synchronize()
        close(d.to_ncdhw(), u.grad, what="upsample bwd %s" % (shape,))
        ops.upsample_bwd(gp, d, accumulate=True)
        torch.cuda.synchronize()
        close(d.to_ncdhw(), 2 * u.grad, tol=2 * BF16_ULP, what="upsample bwd acc")


FUSE_MMA_CASES = [
    # (N, C, full grid (Z, Y, X), low grids, n_same): Y % 16 == 0, Yl <= 32 -> fuse_mma.cu (y interpolation on the tensor cores)
    (2, 16, (4, 64, 40), [(2, 32, 20), (1, 16, 10), (1, 8, 5)], 1),   # the full-resolution exchange: ratios 2 / 4 / 8
    (1, 24, (4, 64, 40), [(2, 32, 20), (1, 16, 10)], 2),              # two same-resolution terms, three channel chunks
    (2, 8, (2, 32, 24), [(1, 16, 12)], 3),                            # half resolution, one low term, three same terms
    (1, 8, (3, 48, 17), [(2, 20, 9), (1, 7, 3)], 1),                  # odd extents, non-integer ratios, three row blocks
    (1, 8, (2, 16, 70), [(1, 8, 35), (1, 1, 1)], 1),                  # long x extent (several units per plane), a 1x1x1 term
]


@pytest.mark.parametrize("case", FUSE_MMA_CASES, ids=str)
def test_fuse_sum_tensor_core_path(case):
    """rtp_fuse_sum with the y interpolation on mma.sync (csrc/fuse_mma.cu, opt-in) against F.interpolate (trilinear,
    align_corners=True) + sum + bias + ReLU in fp32; run-to-run identical; the CUDA-core kernels agree to bf16 round-off."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, C, hi, lows, n_same = case
    same = [rnd(N, C, *hi, seed=40 + i) for i in range(n_same)]
    low = [rnd(N, C, *g, seed=50 + i) for i, g in enumerate(lows)]
    bias = rnd(C, seed=60)
    pre = sum(same) + bias.view(1, C, 1, 1, 1)
    for u in low:
        pre = pre + F.interpolate(u, size=hi, mode="trilinear", align_corners=True)
    ref = F.relu(pre)
    import os
    base = ops.fuse_sum(P8(N, C, *hi), [to_p8(t) for t in same], [to_p8(u) for u in low], bias=bias.cuda(), relu=True)
    os.environ["RTP_FUSE_MMA"] = "1"  # the tensor-core kernel is opt-in (measured slower than the tile kernel)
    try:
        out = ops.fuse_sum(P8(N, C, *hi), [to_p8(t) for t in same], [to_p8(u) for u in low], bias=bias.cuda(), relu=True)
        torch.cuda.synchronize()
        close(out.to_ncdhw(), ref, tol=1.5 * BF16_ULP, what="fuse_sum (mma) %s" % (case,))
        first = out.to_ncdhw().clone()
        out2 = ops.fuse_sum(P8(N, C, *hi), [to_p8(t) for t in same], [to_p8(u) for u in low], bias=bias.cuda(), relu=True)
        torch.cuda.synchronize()
        assert torch.equal(first, out2.to_ncdhw())
        # without ReLU / bias
        out3 = ops.fuse_sum(P8(N, C, *hi), [to_p8(t) for t in same], [to_p8(u) for u in low])
        torch.cuda.synchronize()
        close(out3.to_ncdhw(), pre - bias.view(1, C, 1, 1, 1), tol=1.5 * BF16_ULP, what="fuse_sum (mma), plain")
    finally:
        del os.environ["RTP_FUSE_MMA"]
    close(first, base.to_ncdhw(), tol=1.5 * BF16_ULP, what="tensor-core vs CUDA-core fuse_sum")


UPBWD_MMA_CASES = [
    # (N, C, full grid (Z, Y, X), low grid): Y % 16 == 0 and Yl <= 32 -> upsample_mma.cu (y reduction on the tensor cores)
    (2, 16, (4, 64, 40), (2, 32, 20)),   # ratio 2: two 16-row blocks of yl, four k16 steps, banded weights
    (2, 16, (4, 64, 40), (1, 16, 10)),   # ratio 4 (one z plane at low resolution)
    (1, 24, (2, 64, 56), (2, 8, 7)),     # ratio 8, three channel chunks
    (2, 8, (2, 32, 24), (1, 16, 12)),    # half -> quarter resolution: two k16 steps
    (1, 8, (3, 48, 17), (2, 20, 9)),     # odd extents, three k16 steps, non-integer ratio
    (1, 8, (2, 16, 160), (2, 8, 80)),    # long x extent: several row chunks per unit, rolling accumulators across chunks
]


@pytest.mark.parametrize("case", UPBWD_MMA_CASES, ids=str)
def test_upsample_bwd_tensor_core_path(case):
    """rtp_upsample_bwd with the y reduction on mma.sync and the thread-local x reduction (csrc/upsample_mma.cu) against
    autograd of F.interpolate (trilinear, align_corners=True); fixed summation order; accumulate mode."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, C, hi, lo = case
    u = rnd(N, C, *lo, seed=31).requires_grad_(True)
    up = F.interpolate(u, size=hi, mode="trilinear", align_corners=True)
    g = rnd(N, C, *hi, seed=32)
    up.backward(g)
    gp = to_p8(g)
    d = P8(N, C, *lo)
    ops.upsample_bwd(gp, d)
    torch.cuda.synchronize()
    close(d.to_ncdhw(), u.grad, what="upsample bwd (mma) %s -> %s" % (hi, lo))
    first = d.to_ncdhw().clone()
    ops.upsample_bwd(gp, d)  # fixed summation order: run-to-run identical
    torch.cuda.synchronize()
    assert torch.equal(first, d.to_ncdhw())
    ops.upsample_bwd(gp, d, accumulate=True)
    torch.cuda.synchronize()
    close(d.to_ncdhw(), 2 * u.grad, tol=2 * BF16_ULP, what="upsample bwd (mma) accumulate")


def test_batched_weight_packs_match_single_launches():
    """PackedWeights.refresh_async rebuilds every pack with ONE rtp_weight_pack_batch launch: bit-identical to the per-weight
    launches (generic packs of both modes with an input-channel slice, k3s1 packs plain / transposed / windowed)."""
    from rtpose_b200 import ops
    ws = [torch.randn(40, 24, 3, 3, 3, device="cuda"), torch.randn(128, 32, 1, 1, 1, device="cuda"),
          torch.randn(32, 32, 3, 3, 3, device="cuda"), torch.randn(64, 128, 3, 3, 3, device="cuda")]

    def build(pk):
        return [pk.get(ws[0], 0)[0], pk.get(ws[0], 1, ci0=8, ci_n=16)[0], pk.get(ws[1], 0)[0], pk.get(ws[1], 1)[0],
                pk.get_k3s1(ws[2], 32, 32, False), pk.get_k3s1(ws[2], 32, 32, True),
                pk.get_k3s1(ws[3], 64, 32, True, ci_window=(32, 32)), pk.get_k3s1(ws[3], 128, 64, False)]
    pk = ops.PackedWeights()
    packs = build(pk)
    before = [p.clone() for p in packs]
    for w in ws:
        w.mul_(-1.5)  # in place: same storage, new version (what the optimizer does)
    pk.refresh_async()
    got = build(pk)  # joins the pack stream; every entry is current again, nothing is repacked here
    torch.cuda.synchronize()
    assert pk._batch is not None and pk._batch[2] == len(packs)
    assert all(g.data_ptr() == p.data_ptr() for g, p in zip(got, packs))
    ref = build(ops.PackedWeights())  # fresh cache: one launch per pack
    torch.cuda.synchronize()
    for g, r, b in zip(got, ref, before):
        assert torch.equal(g, r)
        assert not torch.equal(g, b)


def test_sparse_regression_head_backward_matches_dense(ctx):
    """rtp_reg_head_bwd_sparse (backward of the regression branch's last conv from the loss gradient, which is non-zero at
    the target voxels only) against the dense kernels on the same tensors: input gradient with the ReLU mask, weight and bias
    gradients; targets on the volume border, duplicate targets, adjacent targets (overlapping neighbourhoods), accumulate."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, R, grid, M = 3, 32, 45, (6, 16, 24), 15
    Z, Y, X = grid
    g = torch.Generator().manual_seed(9)
    t_in = bf(torch.randn(N, Cin, *grid, generator=g))           # hidden activations (sign = ReLU gate)
    w = bf(torch.randn(R, Cin, 3, 3, 3, generator=g) * 0.05).cuda()
    zz = torch.randint(0, Z, (N, M), generator=g); yy = torch.randint(0, Y, (N, M), generator=g); xx = torch.randint(0, X, (N, M), generator=g)
    zz[0, 0], yy[0, 0], xx[0, 0] = 0, 0, 0                        # corner
    zz[0, 1], yy[0, 1], xx[0, 1] = Z - 1, Y - 1, X - 1            # opposite corner
    zz[1, 1], yy[1, 1], xx[1, 1] = zz[1, 0], yy[1, 0], xx[1, 0]   # duplicate voxel
    zz[2, 1], yy[2, 1], xx[2, 1] = zz[2, 0], yy[2, 0], min(int(xx[2, 0]) + 1, X - 1)  # neighbours
    ind = (zz * Y * X + yy * X + xx).to(torch.int64)
    dyd = torch.zeros(N, R, *grid)
    vals = bf(torch.randn(N, M, R, generator=g))
    for n in range(N):
        for j in range(M):
            dyd[n, :, zz[n, j], yy[n, j], xx[n, j]] = vals[n, j]    # duplicates: last write wins (the tensor is the source of truth)
    dyp, tp = to_p8(dyd), to_p8(t_in)
    # dense reference with the repo's own kernels
    dW_d, db_d = torch.zeros(R, Cin, 3, 3, 3, device="cuda"), torch.zeros(R, device="cuda")
    ops.conv_wgrad(tp, dyp, 3, 1, dW_d)
    ops.channel_sum(dyp, db_d)
    dt_d = P8(N, Cin, *grid)
    ops.conv_dgrad(ctx, dyp, w, 1, dt_d, mask=tp)
    # sparse
    dW_s, db_s = torch.full((R, Cin, 3, 3, 3), 7.0, device="cuda"), torch.full((R,), 7.0, device="cuda")
    wide = P8(N, 2 * Cin, *grid)
    wide.buf.fill_(1.0)                                           # garbage that the zero-fill must remove
    dt_s = wide.channels(Cin, Cin)                                # a channel view, like the merged head gradient
    ops.reg_head_bwd_sparse(dyp, tp, ind.cuda(), w, dt_s, dW_s, False, db_s, False)
    torch.cuda.synchronize()
    close(dt_s.to_ncdhw(), dt_d.to_ncdhw(), tol=1e-2, what="sparse vs dense dgrad of the regression head")
    close(dW_s, dW_d, tol=1e-3, what="sparse vs dense weight gradient")
    close(db_s, db_d, tol=1e-4, what="sparse vs dense bias gradient")
    # against fp32 autograd as well
    xr = t_in.clone().requires_grad_(True)
    wr = w.cpu().clone().requires_grad_(True)
    br = torch.zeros(R, requires_grad=True)
    F.conv3d(xr, wr, br, padding=1).backward(dyd)
    close(dt_s.to_ncdhw(), xr.grad * (t_in > 0), what="sparse dgrad vs autograd")
    close(dW_s, wr.grad, tol=1e-3, what="sparse wgrad vs autograd")
    close(db_s, br.grad, tol=1e-4, what="sparse bias gradient vs autograd")
    ops.reg_head_bwd_sparse(dyp, tp, ind.cuda(), w, dt_s, dW_s, True, db_s, True)
    torch.cuda.synchronize()
    close(dW_s, 2 * wr.grad, tol=1e-3, what="sparse wgrad accumulate")
    close(db_s, 2 * br.grad, tol=1e-4, what="sparse bias accumulate")


def test_unit_lists_for_sparse_operands(ctx):
    """rtp_active_units + the unit-list variants of conv_k3s1 (dgrad, accumulate) and wgrad_k3s1: with a gradient that is
    non-zero only within one voxel of a few target voxels, visiting only the listed (sample, tile) units gives the dense
    result — the dgrad needs the tiles within 2 voxels (x, y) of a target, the weight gradient those within 1."""
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8
    N, C, Cx, grid, M = 4, 32, 128, (6, 64, 160), 5
    Z, Y, X = grid
    g = torch.Generator().manual_seed(13)
    zz = torch.randint(0, Z, (N, M), generator=g); yy = torch.randint(0, Y, (N, M), generator=g); xx = torch.randint(0, X, (N, M), generator=g)
    zz[0, 0], yy[0, 0], xx[0, 0] = 0, 0, 0
    zz[1, 0], yy[1, 0], xx[1, 0] = Z - 1, Y - 1, X - 1
    ind = (zz * Y * X + yy * X + xx).to(torch.int64).cuda()
    dy = torch.zeros(N, C, *grid)
    for n in range(N):
        for j in range(M):
            z0, z1 = max(0, int(zz[n, j]) - 1), min(Z, int(zz[n, j]) + 2)
            y0, y1 = max(0, int(yy[n, j]) - 1), min(Y, int(yy[n, j]) + 2)
            x0, x1 = max(0, int(xx[n, j]) - 1), min(X, int(xx[n, j]) + 2)
            dy[n, :, z0:z1, y0:y1, x0:x1] = bf(torch.randn(C, z1 - z0, y1 - y0, x1 - x0, generator=g))
    dyp = to_p8(dy)
    xin = to_p8(rnd(N, Cx, *grid, seed=14))
    w = bf(torch.randn(C, Cx, 3, 3, 3, generator=g) * 0.05).cuda()   # conv Cx -> C; dgrad: dy (C) -> dx (Cx)
    like = P8(N, Cx, *grid)
    u1 = ops.active_units(ind, like, 1, "t1")
    u2 = ops.active_units(ind, like, 2, "t2")
    torch.cuda.synchronize()
    ntile = (X * (Y + 2) + 127) // 128
    n1, n2 = int(u1[1][0]), int(u2[1][0])
    assert 0 < n1 <= n2 < N * ntile // 2, (n1, n2, N * ntile)
    l2 = u2[0][:n2].cpu()
    assert bool((l2[1:] > l2[:-1]).all())
    # every voxel within 2 of a target lies in a listed unit
    Yp = Y + 2
    listed = set(l2.tolist())
    for n in range(N):
        for j in range(M):
            for dxv in range(-2, 3):
                for dyv in range(-2, 3):
                    x, y = int(xx[n, j]) + dxv, int(yy[n, j]) + dyv
                    if 0 <= x < X and 0 <= y < Y:
                        assert n * ntile + ((x + 1) * Yp + (y + 1) - Yp) // 128 in listed
    dense = P8(N, Cx, *grid)
    ops.conv_dgrad(ctx, dyp, w, 1, dense)
    base = rnd(N, Cx, *grid, seed=15)
    viau = to_p8(base)
    ops.conv_dgrad(ctx, dyp, w, 1, viau, accumulate=True, units=u2)
    torch.cuda.synchronize()
    close(viau.to_ncdhw(), bf(base).cuda() + dense.to_ncdhw(), tol=2 * BF16_ULP, what="dgrad over the listed units (accumulate)")
    far = (dense.to_ncdhw() == 0)
    assert torch.equal(viau.to_ncdhw()[far], bf(base).cuda()[far])  # untouched where the dense result is exactly zero
    dW_d, dW_u = torch.zeros(C, Cx, 3, 3, 3, device="cuda"), torch.zeros(C, Cx, 3, 3, 3, device="cuda")
    ops.conv_wgrad(xin, dyp, 3, 1, dW_d)
    lib.call_counts.clear()
    ops.conv_wgrad(xin, dyp, 3, 1, dW_u, units=u1)
    torch.cuda.synchronize()
    assert lib.call_counts.get("rtp_wgrad_k3s1_units", 0) == Cx // 32
    close(dW_u, dW_d, tol=1e-5, what="weight gradient over the listed units")


def test_grad_add_channel_sum_stem():
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8, _stream
    N, C, grid = 2, 32, (3, 6, 8)
    s, m, d0 = rnd(N, C, *grid, seed=22), rnd(N, C, *grid, seed=23), rnd(N, C, *grid, seed=24)
    d = to_p8(d0)
    ops.grad_add(to_p8(s), d, mask=to_p8(m), accumulate=True)
    out = torch.zeros(C, device="cuda")
    ops.channel_sum(to_p8(s), out)
    torch.cuda.synchronize()
    close(d.to_ncdhw(), d0 + s * (m > 0), what="grad_add")
    close(out, s.sum((0, 2, 3, 4)), tol=1e-4, what="channel_sum")
    # 1 -> C stem
    x = rnd(N, 1, *grid, seed=25)
    w, b = rnd(C, 1, 1, 1, 1, seed=26).requires_grad_(True), rnd(C, seed=27).requires_grad_(True)
    y = F.conv3d(x, w, b)
    dy = rnd(*y.shape, seed=28)
    y.backward(dy)
    xp, yp = to_p8(x), P8(N, C, *grid)
    wd, bd = w.detach().reshape(-1).cuda(), b.detach().cuda()  # keep alive: raw pointers are passed below
    lib.call("rtp_stem_fwd", xp.struct(), wd.data_ptr(), bd.data_ptr(), C, yp.struct(), _stream())
    gw, gb = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dyp = to_p8(dy)
    lib.call("rtp_stem_bwd", xp.struct(), dyp.struct(), C, gw.data_ptr(), gb.data_ptr(), 0,
             ops.gn_ws(yp).data_ptr(), _stream())
    torch.cuda.synchronize()
    close(yp.to_ncdhw(), y, what="stem fwd")
    close(gw, w.grad.reshape(-1), tol=1e-3, what="stem dw")
    close(gb, b.grad, tol=1e-3, what="stem db")


def test_ingest_matches_oracle():
    from oracle import hrpose_oracle as O
    from rtpose_b200 import lib
    from rtpose_b200.p8 import P8, _stream
    rs = np.random.RandomState(3)
    raw = rs.uniform(-2, 12, size=(2, 16, 32, 128, 256)).astype(np.float16)
    ref = np.stack([O.ingest_cube(raw[n], (0.0, 10.0)) for n in range(2)])
    rawd = torch.from_numpy(raw).cuda()
    keep = []
    dst = P8(2, 16, 16, 64, 160)
    f32 = torch.empty((2, 16, 16, 64, 160), dtype=torch.float32, device="cuda")
    z0, _, y0, _, x0, _ = O.ROI_IDX
    lib.call("rtp_ingest_pack", rawd.data_ptr(), 2, 16, 32, 128, 256, z0, y0, x0, 0.0, 10.0, 1, dst.struct(),
             f32.data_ptr(), _stream())
    torch.cuda.synchronize()
    assert np.array_equal(f32.cpu().numpy(), ref), "fp32 side output must be bit-identical to numpy"
    assert torch.equal(dst.to_ncdhw().cpu(), bf(torch.from_numpy(ref)))
    # without the fp32 side output the vectorised kernel runs (16-byte loads of the aligned superset of each row): same bits
    dstv = P8(2, 16, 16, 64, 160)
    lib.call("rtp_ingest_pack", rawd.data_ptr(), 2, 16, 32, 128, 256, z0, y0, x0, 0.0, 10.0, 1, dstv.struct(), None, _stream())
    torch.cuda.synchronize()
    assert torch.equal(dstv.buf, dst.buf), "vectorised ingest differs from the scalar kernel (pads included)"
    # single-channel 'zyx_real' cube and the crop-only phase variant
    one = rs.uniform(25000, 60000, size=(1, 1, 32, 128, 256)).astype(np.float16)  # finite in fp16 (150000 is not)
    d1 = P8(1, 1, 16, 64, 160)
    keep.append(torch.from_numpy(one).cuda())
    lib.call("rtp_ingest_pack", keep[-1].data_ptr(), 1, 1, 32, 128, 256, z0, y0, x0, 30000.0,
             20000.0, 1, d1.struct(), None, _stream())
    torch.cuda.synchronize()
    assert torch.equal(d1.to_ncdhw().cpu(), bf(torch.from_numpy(O.ingest_cube(one[0, 0], (30000.0, 50000.0))[None])))
    ph = rs.uniform(-1, 1, size=(1, 2, 4, 32, 128, 256)).astype(np.float16)
    d2 = P8(1, 8, 16, 64, 160)
    keep.append(torch.from_numpy(ph).cuda())
    lib.call("rtp_ingest_pack", keep[-1].data_ptr(), 1, 8, 32, 128, 256, z0, y0, x0, 0.0, 1.0, 0,
             d2.struct(), None, _stream())
    torch.cuda.synchronize()
    assert torch.equal(d2.to_ncdhw().cpu(), bf(torch.from_numpy(O.ingest_cube_phase(ph[0])[None])))


@pytest.mark.parametrize("one_hm", [True, False])
def test_head_loss_and_decode(one_hm):
    from oracle import hrpose_oracle as O
    from rtpose_b200.engine import Engine
    grid, N = (8, 16, 24), 3
    ncls, R = (1, 45) if one_hm else (15, 3)
    rs = np.random.RandomState(5)
    poses = [O.synth_pose(rs, grid) for _ in range(N)]
    tgt = O.batch_targets(poses, grid, one_hm)
    hm = (rnd(N, ncls, *grid, seed=30) * 0.5 - 2.0).requires_grad_(True)
    hm.data = bf(hm.data)
    reg = rnd(N, R, *grid, seed=31).requires_grad_(True)
    cw = [1.0] * 45 if one_hm else [1.0, 1.5, 2.0]
    L = O.head_loss({"hm": hm, "reg": reg}, tgt, 0.5, cw)
    L["loss"].backward()
    eng = Engine("hr_tiny_feat32_zyx_l4_in32", "top", {}, R, ncls, 0.5, cw)
    eng.begin()
    hp, rp = to_p8(hm.detach()), to_p8(reg.detach())
    out = eng.loss(hp, rp, tgt["hm"].cuda(), tgt["ind"].cuda(), tgt["mask"].cuda(), tgt["cat"].cuda(),
                   tgt["anno_pose"].cuda())
    torch.cuda.synchronize()
    o = out.cpu()
    assert abs(o[0].item() - L["loss"].item()) <= 1e-4 * abs(L["loss"].item())
    assert abs(o[1].item() - L["hm_loss"].item()) <= 1e-4 * abs(L["hm_loss"].item())
    assert abs(o[2].item() - L["loc_loss"].item()) <= 1e-4 * abs(L["loc_loss"].item()) + 1e-6
    assert o[3].item() == L["num_positive"].item()
    np.testing.assert_allclose(o[4:].numpy(), L["loc_loss_elem"].detach().numpy(), rtol=1e-4, atol=1e-6)
    close(hp.grad.to_ncdhw(), hm.grad, tol=2 * BF16_ULP, what="d loss / d hm")
    close(rp.grad.to_ncdhw(), reg.grad, tol=2 * BF16_ULP, what="d loss / d reg")
    # decode: indices bit-exact against the oracle run on the same bf16 heatmap / regression maps
    idx, score, xyz = eng.decode(hp, to_p8(bf(reg.detach())), O.VOXEL_SIZE, O.PC_RANGE)
    torch.cuda.synchronize()
    kps, ref_idx = O.decode(hm.detach(), bf(reg.detach()))
    assert idx.cpu().tolist() == ref_idx
    for n in range(N):
        for j, kp in enumerate(kps[n]):
            c = kp[0] if not one_hm else 0
            got = xyz[n, c, 3 * (j if one_hm else 0):3 * (j if one_hm else 0) + 3].cpu().numpy()
            np.testing.assert_allclose(got, np.array(kp[1:4], dtype=np.float32), rtol=1e-6, atol=1e-6)
            assert abs(score[n, c].item() - kp[4]) <= 1e-6


def test_head_loss_flags_sparse_regression_gradient_and_spare_chunk():
    """rtp_head_loss_flags: with RTP_LOSS_SPARSE_DREG the regression gradient is cleared + accumulated at the target voxels only
    (identical there to the dense call, untouched elsewhere), and chunks of d_hm behind the class chunks are zero-filled."""
    from oracle import hrpose_oracle as O
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8
    grid, N, ncls, R = (8, 16, 24), 3, 15, 3
    rs = np.random.RandomState(6)
    tgt = O.batch_targets([O.synth_pose(rs, grid) for _ in range(N)], grid, False)
    tgt = {k: v.cuda() for k, v in tgt.items()}
    tgt["ind"][1, 3] = tgt["ind"][1, 2]   # two targets (classes 2 and 3) of a sample on one voxel: their regression gradients add up
    tgt["mask"][1, 2:4] = 1
    tgt["mask"][2, 5] = 0                 # a masked-out target: contributes nothing, its voxel still gets a defined (zero) gradient
    hp, rp = to_p8(bf(rnd(N, ncls, *grid, seed=32) * 0.5 - 2.0)), to_p8(rnd(N, R, *grid, seed=33))
    cw = torch.ones(R, device="cuda")
    M = tgt["ind"].shape[1]
    ws = torch.empty(lib.load().rtp_head_loss_workspace_bytes(N, ncls, *grid), dtype=torch.uint8, device="cuda")

    def run(flags, fill):
        d_hm, d_reg = P8(N, 24, *grid), P8(N, R, *grid)
        d_hm.buf.fill_(fill)
        d_reg.buf.fill_(fill)
        out = torch.empty(4 + R, device="cuda")
        lib.call("rtp_head_loss_flags", hp.struct(), rp.struct(), ncls, R, tgt["hm"].data_ptr(), tgt["ind"].data_ptr(),
                 tgt["mask"].data_ptr(), tgt["cat"].data_ptr(), tgt["anno_pose"].data_ptr(), M, 0.5, cw.data_ptr(), 1.0,
                 out.data_ptr(), d_hm.struct(), d_reg.struct(), flags, ws.data_ptr(), ops._stream())
        torch.cuda.synchronize()
        return out.cpu(), d_hm.to_ncdhw().cpu(), d_reg.to_ncdhw().cpu()

    o_d, hm_d, reg_d = run(0, 7.0)
    o_s, hm_s, reg_s = run(lib.RTP_LOSS_SPARSE_DREG, 7.0)
    assert torch.equal(o_d, o_s)
    assert torch.equal(hm_d, hm_s)
    assert float(hm_d[:, 15:].abs().max()) == 0.0         # class padding (channel 15) and the spare chunk (16..23)
    assert float(hm_d[:, :15].abs().max()) > 0.0
    Z, Y, X = grid
    at_target = torch.zeros(N, Z * Y * X, dtype=torch.bool)
    at_target.scatter_(1, tgt["ind"].cpu(), True)
    at_target = at_target.view(N, 1, Z, Y, X).expand(N, R, Z, Y, X)
    assert torch.equal(reg_s[at_target], reg_d[at_target])
    assert float(reg_d[at_target].abs().max()) > 0.0
    assert bool((reg_s[~at_target] == 7.0).all())          # untouched away from the targets
    assert float(reg_d[~at_target].abs().max()) == 0.0     # the dense call zero-fills


def test_decode_tie_breaks_to_lowest_reference_index():
    from rtpose_b200.engine import Engine
    grid, N = (4, 6, 10), 2
    hm = torch.full((N, 1, *grid), -3.0)
    Z, Y, X = grid
    # equal maxima at several voxels: reference flat index z*Y*X + y*X + x must pick the lowest
    for (z, y, x) in [(3, 1, 2), (1, 5, 9), (1, 2, 7), (2, 0, 0)]:
        hm[0, 0, z, y, x] = 1.5
    hm[1, 0, 0, 0, 3] = 0.25
    hm[1, 0, 0, 0, 4] = 0.25
    reg = torch.zeros(N, 45, *grid)
    eng = Engine("hr_tiny_feat32_zyx_l4_in32", "top", {}, 45, 1, 0.5, [1.0] * 45)
    idx, _, _ = eng.decode(to_p8(hm), to_p8(reg), (1, 1, 1), (0, 0, 0))
    torch.cuda.synchronize()
    assert idx.cpu().flatten().tolist() == [1 * Y * X + 2 * X + 7, 3]
    assert idx.cpu().flatten().tolist() == [int(torch.argmax(torch.sigmoid(hm[n, 0]).flatten())) for n in range(N)]


# ------------------------------------------------------------------------------------------------ plane-streaming conv
K3S1_CASES = [
    # N, Cin, Cout, grid(Z,Y,X), bias, relu, res
    (2, 32, 32, (16, 30, 20), False, True, True),      # all 16 planes resident, several tiles
    (1, 32, 32, (16, 64, 160), False, True, False),    # the full-resolution shape (83 tiles)
    (16, 32, 32, (8, 30, 40), False, False, False),    # 160 units > 148 SMs: persistent loop, block recycling
    (1, 128, 32, (16, 14, 12), True, True, False),     # wide K: 4 passes, double-buffered weights
    (2, 32, 45, (16, 10, 12), True, False, False),     # NPo = 48 -> z-chunks of 10 + 6 planes
    (1, 64, 64, (12, 10, 14), False, False, False),    # NPo = 64 -> z-chunks of 8 + 4
    (2, 32, 1, (16, 12, 10), True, False, False),      # NPo = 16
    (1, 16, 32, (3, 9, 7), False, True, False),        # K = 16, odd sizes
    (1, 64, 32, (1, 8, 8), False, False, False),       # single plane
]


@pytest.mark.parametrize("case", K3S1_CASES, ids=[str(c) for c in K3S1_CASES])
def test_conv_k3s1_forward_and_dgrad(ctx, case):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid, use_bias, relu, use_res = case
    x = rnd(N, Cin, *grid, seed=40).requires_grad_(True)
    w = rnd(Cout, Cin, 3, 3, 3, seed=41, scale=(Cin * 27) ** -0.5).requires_grad_(True)
    b = rnd(Cout, seed=42) if use_bias else None
    pre = F.conv3d(x, w, b, padding=1)
    res = rnd(*pre.shape, seed=43) if use_res else None
    ref = pre + res if use_res else pre
    ref = F.relu(ref) if relu else ref
    xp = to_p8(x.detach())
    assert ops.k3s1_eligible(xp, (Cin + 15) // 16 * 16, (Cout + 15) // 16 * 16), "case should take the fast path"
    out = P8(N, Cout, *grid)
    wc = w.detach().cuda()
    ops.conv_forward(ctx, xp, wc, 1, out, bias=b.cuda() if use_bias else None, relu=relu,
                     res=to_p8(res) if use_res else None)
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, what="k3s1 fwd")
    # pads must stay zero
    full = out.buf[out.offset:out.offset + out.N * out.n_stride].view(out.N, out.C8, out.Z, out.X + 2, out.Y + 2, 8).float()
    assert full[:, :, :, 0].abs().max().item() == 0 and full[:, :, :, -1].abs().max().item() == 0
    assert full[:, :, :, :, 0].abs().max().item() == 0 and full[:, :, :, :, -1].abs().max().item() == 0
    # dgrad through the same kernel (flipped / transposed pack), with mask + accumulate
    dy = rnd(*pre.shape, seed=44)
    pre.backward(dy)
    dyp = to_p8(dy)
    if ops.k3s1_eligible(dyp, (Cout + 15) // 16 * 16, (Cin + 15) // 16 * 16):
        dx = P8(N, Cin, *grid)
        ops.conv_dgrad(ctx, dyp, wc, 1, dx)
        torch.cuda.synchronize()
        close(dx.to_ncdhw(), x.grad, what="k3s1 dgrad")
        mask, base = rnd(N, Cin, *grid, seed=45), rnd(N, Cin, *grid, seed=46)
        dx2 = to_p8(base)
        ops.conv_dgrad(ctx, dyp, wc, 1, dx2, mask=to_p8(mask), accumulate=True)
        torch.cuda.synchronize()
        close(dx2.to_ncdhw(), base + x.grad * (mask > 0), tol=2 * BF16_ULP, what="k3s1 dgrad mask+acc")


@pytest.mark.parametrize("chan", [64, 32], ids=["single_lane_64", "dual_lane_32"])
def test_conv_k3s1_tail_split_half_depth_units(ctx, chan):
    """conv_k3s1 whose unit count leaves a partly filled last round (the half-resolution convs of the bench: 16 samples x 22
    tiles = 352 units on 148 SMs): the units of that round run as two half units each — half the output planes (single lane,
    64 channels) or one z-chunk lane per CTA (dual lane, 32 channels).  Forward (bias, ReLU, residual) and dgrad (mask,
    accumulate) against torch on the GPU; the fused GroupNorm statistics of the dual-lane kernel against a separate pass."""
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid = 16, chan, chan, (8, 32, 80)
    nsm = ops.num_sms()
    ntile = ((grid[2]) * (grid[1] + 2) + 127) // 128
    assert N * ntile > nsm and 0 < (N * ntile) % nsm <= nsm // 2, "shape no longer produces a short last round on this GPU"
    g = torch.Generator(device="cuda").manual_seed(3)
    x = bf(torch.randn(N, Cin, *grid, device="cuda", generator=g)).requires_grad_(True)
    w = bf(torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) * (Cin * 27) ** -0.5).requires_grad_(True)
    b = bf(torch.randn(Cout, device="cuda", generator=g))
    res = bf(torch.randn(N, Cout, *grid, device="cuda", generator=g))
    pre = F.conv3d(x, w, b, padding=1)
    ref = F.relu(pre + res)
    xp = P8.from_ncdhw(x.detach())
    out = P8(N, Cout, *grid)
    ops.conv_forward(ctx, xp, w.detach(), 1, out, bias=b, relu=True, res=P8.from_ncdhw(res))
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, what="k3s1 fwd with half-depth tail units")
    dy = bf(torch.randn(N, Cout, *grid, device="cuda", generator=g))
    pre.backward(dy)
    mask = bf(torch.randn(N, Cin, *grid, device="cuda", generator=g))
    base = bf(torch.randn(N, Cin, *grid, device="cuda", generator=g))
    dx = P8.from_ncdhw(base)
    ops.conv_dgrad(ctx, P8.from_ncdhw(dy), w.detach(), 1, dx, mask=P8.from_ncdhw(mask), accumulate=True)
    torch.cuda.synchronize()
    close(dx.to_ncdhw(), base + x.grad * (mask > 0), tol=2 * BF16_ULP, what="k3s1 dgrad with half-depth tail units")
    if chan == 32 and ops.stat_fusable(xp, w.detach(), False):
        out2 = P8(N, Cout, *grid)
        _, st = ops.conv_forward(ctx, xp, w.detach(), 1, out2, bias=b, relu=True, res=P8.from_ncdhw(res), stat=("stats", 8, 1e-5))
        want = ops.gn_stats(out2, 8)
        torch.cuda.synchronize()
        assert torch.equal(out2.to_ncdhw(), out.to_ncdhw())
        close(st, want, tol=1e-4, what="fused statistics with split tail units")


def test_generic_path_still_used_when_fast_path_disabled(ctx):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    x, w = rnd(1, 32, 4, 10, 12, seed=47), rnd(32, 32, 3, 3, 3, seed=48, scale=0.03)
    ref = F.conv3d(x, w, padding=1)
    ops.USE_K3S1 = False
    try:
        out = ops.conv_forward(ctx, to_p8(x), w.cuda(), 1, P8(1, 32, 4, 10, 12))
    finally:
        ops.USE_K3S1 = True
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, what="generic conv")


WGRAD3_CASES = [
    # N, Cin, Cout, grid(Z,Y,X)
    (2, 32, 32, (16, 30, 20)),
    (1, 32, 32, (16, 64, 160)),     # full resolution: 83 tiles, last one partial
    (16, 32, 32, (4, 30, 40)),      # 160 units > 148 SMs
    (2, 32, 45, (8, 10, 12)),       # NP = 48
    (2, 32, 1, (8, 12, 10)),        # NP = 16, dY has one chunk
    (1, 128, 64, (6, 14, 12)),      # 4 input groups x 2 output groups (the merged head conv)
    (1, 64, 64, (3, 8, 10)),
    (1, 32, 45, (2, 70, 9)),        # span mode (row >= 65 positions), NP = 48, dY ring wraps with two planes
    (2, 32, 32, (1, 64, 5)),        # span mode, single plane
    (1, 32, 32, (1, 9, 7)),         # copy mode, single plane
]


@pytest.mark.parametrize("case", WGRAD3_CASES, ids=[str(c) for c in WGRAD3_CASES])
def test_wgrad_k3s1(case):
    from rtpose_b200 import lib, ops
    N, Cin, Cout, grid = case
    x = rnd(N, Cin, *grid, seed=50)
    dy = rnd(N, Cout, *grid, seed=51)
    w = torch.zeros(Cout, Cin, 3, 3, 3, requires_grad=True)
    F.conv3d(x, w, padding=1).backward(dy)
    assert lib.load().rtp_wgrad_k3s1_supported(32, 32, grid[0], grid[2], grid[1])
    dW = torch.full(w.shape, 3.0, device="cuda")
    xp, dyp = to_p8(x), to_p8(dy)
    ops.conv_wgrad(xp, dyp, 3, 1, dW)
    torch.cuda.synchronize()
    close(dW, w.grad, tol=2e-3, what="wgrad k3s1")
    ops.conv_wgrad(xp, dyp, 3, 1, dW, accumulate=True)
    torch.cuda.synchronize()
    close(dW, 2 * w.grad, tol=2e-3, what="wgrad k3s1 accumulate")
    # generic kernel agrees
    ops.USE_WGRAD_K3S1 = False
    try:
        dW2 = torch.zeros_like(dW)
        ops.conv_wgrad(xp, dyp, 3, 1, dW2)
    finally:
        ops.USE_WGRAD_K3S1 = True
    torch.cuda.synchronize()
    close(dW2, w.grad, tol=2e-3, what="wgrad generic")


STAT_CASES = [
    # N, Cin, Cout, grid(Z,Y,X)
    (2, 32, 32, (16, 30, 20)),
    (3, 32, 32, (5, 64, 37)),      # several units per CTA across samples, z-chunking
    (16, 32, 32, (4, 20, 24)),     # more samples than the common case: slab rows for every n
    (2, 64, 32, (6, 12, 10)),      # two K passes
    (2, 32, 16, (4, 10, 12)),      # 16 result channels
    (16, 32, 32, (16, 64, 160)),   # the BASELINE shape: 9 units per CTA, every sample touched by several CTAs
]


@pytest.mark.parametrize("case", STAT_CASES, ids=[str(c) for c in STAT_CASES])
def test_conv_k3s1_fused_groupnorm_statistics(ctx, case):
    """stat_mode 1 / 2 of rtp_conv_k3s1: the statistics that come out of the conv epilogue equal the ones the separate
    reduction kernels compute from the stored tensor (different summation order: fp32 round-off only)."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid = case
    x = rnd(N, Cin, *grid, seed=70)
    w = rnd(Cout, Cin, 3, 3, 3, seed=71, scale=0.1).cuda()
    res = rnd(N, Cout, *grid, seed=72)
    xp, resp = to_p8(x), to_p8(res)
    assert ops.stat_fusable(xp, w, False)
    # forward: y = relu(conv(x) + res), statistics of y
    y0 = ops.conv_forward(ctx, xp, w, 1, P8(N, Cout, *grid), relu=True, res=resp)
    G = 8
    ref_stats = ops.gn_stats(y0, G)
    y1, stats = ops.conv_forward(ctx, xp, w, 1, P8(N, Cout, *grid), relu=True, res=resp, stat=("stats", G, 1e-5))
    torch.cuda.synchronize()
    assert torch.equal(y0.to_ncdhw(), y1.to_ncdhw())
    torch.testing.assert_close(stats, ref_stats, rtol=2e-5, atol=2e-6)
    # dgrad: dxn = conv_transpose(dy), reductions against the GroupNorm input xg with its statistics
    dy = to_p8(rnd(N, Cout, *grid, seed=73))
    xg = to_p8(rnd(N, Cin, *grid, seed=74) + 0.3)
    assert ops.stat_fusable(dy, w, True) == (Cin <= 32)
    Gx = 8
    st_x = ops.gn_stats(xg, Gx)
    d0 = ops.conv_dgrad(ctx, dy, w, 1, P8(N, Cin, *grid))
    red_ref = torch.empty((N, Cin, 2), dtype=torch.float32, device="cuda")
    from rtpose_b200 import lib
    from rtpose_b200.p8 import _stream
    lib.call("rtp_gn_bwd_reduce", xg.struct(), d0.struct(), Cin, Gx, st_x.data_ptr(), red_ref.data_ptr(),
             ops.gn_ws(xg).data_ptr(), _stream())
    if Cin <= 32:
        d1, red = ops.conv_dgrad(ctx, dy, w, 1, P8(N, Cin, *grid), stat=("red", Gx, xg, st_x))
        torch.cuda.synchronize()
        assert torch.equal(d0.to_ncdhw(), d1.to_ncdhw())
        scale = red_ref.abs().max().item()
        assert (red - red_ref).abs().max().item() <= 2e-5 * scale + 1e-4


S2D_CASES = [
    # N, Cin, Cout, grid(Z,Y,X) (even)
    (2, 32, 32, (8, 20, 24)),
    (1, 32, 64, (4, 12, 16)),
    (2, 32, 32, (2, 64, 10)),
]


@pytest.mark.parametrize("case", S2D_CASES, ids=[str(c) for c in S2D_CASES])
def test_stride2_conv_through_space_to_depth_view(ctx, case):
    """GroupNorm -> stride-2 3x3x3 conv, forward / dgrad / wgrad / GroupNorm backward, computed through the s2d view with
    the stride-1 plane-streaming kernels, against torch (F.group_norm + F.conv3d stride 2) on the same bf16 inputs."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid = case
    x = rnd(N, Cin, *grid, seed=80).requires_grad_(True)
    w = rnd(Cout, Cin, 3, 3, 3, seed=81, scale=0.1).requires_grad_(True)
    gamma = (1.0 + 0.1 * rnd(Cin, seed=82)).requires_grad_(True)
    beta = (0.1 * rnd(Cin, seed=83)).requires_grad_(True)
    xn_ref = F.group_norm(x, 8, gamma, beta, eps=1e-5)
    xn_ref.retain_grad()
    y_ref = F.conv3d(bf(xn_ref.detach()).requires_grad_(False) + (xn_ref - xn_ref.detach()), w, stride=2, padding=1)
    og = tuple((g - 1) // 2 + 1 for g in grid)
    dy = rnd(N, Cout, *og, seed=84)
    y_ref.backward(dy)

    xp = to_p8(x.detach())
    wc, gc, bc = w.detach().cuda(), gamma.detach().cuda(), beta.detach().cuda()
    stats = ops.gn_stats(xp, 8)
    xs = ops.gn_apply_s2d(xp, 8, stats, gc, bc, P8(N, 8 * Cin, grid[0] // 2, grid[1] // 2, grid[2] // 2))
    we = ops.s2d_expand(wc)
    ops.S2D_MIN_VOXELS = 0
    assert ops.s2d_eligible(xp, wc)
    y = ops.conv_forward(ctx, xs, we, 1, P8(N, Cout, *og), key=("t", case), version=0)
    ym = ops.conv_forward(ctx, xs, we, 1, P8(N, Cout, *og), key=("t", case), version=0,
                          tap_mask=[ops.s2d_tap_mask(par, False) for par in range(8)])
    torch.cuda.synchronize()
    close(y.to_ncdhw(), y_ref, what="s2d forward")
    close(ym.to_ncdhw(), y_ref, what="s2d forward, all-zero taps skipped")
    # backward
    dyp = to_p8(dy)
    gwe = torch.zeros_like(we)
    ops.conv_wgrad(xs, dyp, 3, 1, gwe)
    gw = torch.full(wc.shape, 7.0, device="cuda")
    ops.s2d_fold(gwe, gw, False)
    torch.cuda.synchronize()
    close(gw, w.grad, tol=3e-3, what="s2d wgrad")
    gw2 = torch.full(wc.shape, 7.0, device="cuda")
    ops.conv_wgrad_s2d(xs, dyp, Cin, gw2)            # generic kernel over the view, 27 (parity, offset) taps
    ops.conv_wgrad_s2d(xs, dyp, Cin, gw2, accumulate=True)
    torch.cuda.synchronize()
    close(gw2, 2 * w.grad, tol=3e-3, what="s2d wgrad (plane-streaming kernel over the view when Cin = 32, else gather kernel)")
    old = ops.USE_WGRAD_S2D
    ops.USE_WGRAD_S2D = False
    try:
        gw3 = torch.full(wc.shape, 7.0, device="cuda")
        ops.conv_wgrad_s2d(xs, dyp, Cin, gw3)        # generic kernel over the view, 27 (parity, offset) taps
    finally:
        ops.USE_WGRAD_S2D = old
    torch.cuda.synchronize()
    close(gw3, w.grad, tol=3e-3, what="s2d wgrad (gather kernel over the view)")
    gw4 = torch.full(wc.shape, 7.0, device="cuda")
    ops.conv_wgrad_s2d(xs, dyp, Cin, gw4)
    torch.cuda.synchronize()
    close(gw4, gw3, tol=2e-5, what="s2d wgrad: plane-streaming vs gather kernel (same bf16 operands, fp32 accumulation)")
    dxs = ops.conv_dgrad(ctx, dyp, we, 1, P8(N, 8 * Cin, grid[0] // 2, grid[1] // 2, grid[2] // 2), key=("t", case), version=0,
                         s2d_cin=Cin)
    dg, db = torch.zeros(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
    dx = P8(N, Cin, *grid)
    ops.gn_backward(xp, dxs, 8, stats, gc, dg, db, False, dx, False, s2d=True)
    torch.cuda.synchronize()
    close(dx.to_ncdhw(), x.grad, tol=2e-2, what="s2d dx")
    close(dg, gamma.grad, tol=2e-2, what="s2d dgamma")
    close(db, beta.grad, tol=2e-2, what="s2d dbeta")
    extra = rnd(N, Cin, *grid, seed=77)
    dx5 = P8(N, Cin, *grid)
    ops.gn_backward(xp, dxs, 8, stats, gc, dg, db, False, dx5, False, s2d=True, add=to_p8(extra))
    torch.cuda.synchronize()
    close(dx5.to_ncdhw(), dx.to_ncdhw().cpu() + bf(extra), tol=2 * BF16_ULP, what="s2d dx + add")


@pytest.mark.parametrize("case", [(2, 32, 32, (8, 20, 24)), (1, 32, 32, (2, 64, 10))], ids=str)
def test_sibling_stride2_convs_with_the_groupnorm_affine_folded_in(ctx, case):
    """Two `GroupNorm -> conv3x3x3 stride 2` layers with different (gamma, beta, W) reading the same x (the fuse layers out
    of branch 0, hr3d.py:159-203) computed from ONE space-to-depth view of xhat with the affine folded into each conv
    (csrc/s2d_shared.cu), against torch: outputs (one with ReLU), dW, dgamma, dbeta of both and the summed dL/dx."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid = case
    hg = tuple(g // 2 for g in grid)
    x = rnd(N, Cin, *grid, seed=180).requires_grad_(True)
    ws = [rnd(Cout, Cin, 3, 3, 3, seed=181 + k, scale=0.1).requires_grad_(True) for k in range(2)]
    gammas = [(1.0 + 0.2 * rnd(Cin, seed=183 + k)).requires_grad_(True) for k in range(2)]
    betas = [(0.3 * rnd(Cin, seed=185 + k)).requires_grad_(True) for k in range(2)]
    dys = [rnd(N, Cout, *hg, seed=187 + k) for k in range(2)]
    relus = [True, False]
    y_refs = []
    for k in range(2):
        xn = F.group_norm(x, 8, gammas[k], betas[k], eps=1e-5)
        y = F.conv3d(bf(xn.detach()) + (xn - xn.detach()), ws[k], stride=2, padding=1)
        if relus[k]:
            y = F.relu(y)
        y_refs.append(y)
    sum((y * d).sum() for y, d in zip(y_refs, dys)).backward()

    ops.S2D_MIN_VOXELS = 0
    xp = to_p8(x.detach())
    stats = ops.gn_stats(xp, 8)
    ones, zeros = torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
    V = ops.gn_apply_s2d(xp, 8, stats, ones, zeros, P8(N, 8 * Cin, *hg))
    dV = P8(N, 8 * Cin, *hg)
    masks = [ops.s2d_tap_mask(par, False) for par in range(8)]
    for k in range(2):
        wc, gc, bc = ws[k].detach().cuda(), gammas[k].detach().cuda(), betas[k].detach().cuda()
        we = torch.empty((Cout, 8 * Cin, 3, 3, 3), device="cuda")
        bias_cls = torch.empty((8, Cout), device="cuda")
        ops.s2d_fold_weights(wc, gc, bc, we, bias_cls)
        r1 = P8(1, Cout, *hg)
        ops.s2d_border_bias(bias_cls, r1, Cout)
        rb = P8(N, Cout, *hg, buf=r1.buf, offset=r1.offset, n_stride=0, c_stride=r1.c_stride)
        y = ops.conv_forward(ctx, V, we, 1, P8(N, Cout, *hg), bias=bias_cls[0], relu=relus[k], res=rb, key=("sib", case, k),
                             version=0, tap_mask=masks)
        torch.cuda.synchronize()
        # bias classes against a direct evaluation: conv of the constant beta field with zero padding
        bfield = F.conv3d(betas[k].detach().view(1, Cin, 1, 1, 1).expand(1, Cin, *grid), ws[k].detach(), stride=2, padding=1)[0]
        for cls in range(8):
            z, yy, xx = (0 if cls & 4 else 1), (0 if cls & 1 else 1), (0 if cls & 2 else 1)
            if z >= hg[0] or yy >= hg[1] or xx >= hg[2]:
                continue  # this class has no output position on such a thin grid
            close(bias_cls[cls], bfield[:, z, yy, xx], tol=1e-4, what="border-class bias %d" % cls)
        close(y.to_ncdhw(), y_refs[k], tol=1.5 * BF16_ULP, what="sibling %d forward" % k)
        # backward: dy w.r.t. the pre-ReLU output
        dy = dys[k] * (y_refs[k].detach() > 0).float() if relus[k] else dys[k]
        dyp = to_p8(dy)
        dwp = torch.empty_like(wc)
        ops.conv_wgrad_s2d(V, dyp, Cin, dwp, accumulate=False)
        gw, gg, gb = (torch.full(t.shape, 3.0, device="cuda") for t in (wc, gc, bc))
        ops.s2d_fold_wgrad(dyp, dwp, wc, gc, bc, gw, gg, gb, False, False)
        ops.conv_dgrad(ctx, dyp, we, 1, dV, accumulate=k > 0, key=("sib", case, k), version=0, s2d_cin=Cin)
        torch.cuda.synchronize()
        close(gw, ws[k].grad, tol=5e-3, what="sibling %d dW" % k)
        close(gg, gammas[k].grad, tol=2e-2, what="sibling %d dgamma" % k)
        close(gb, betas[k].grad, tol=2e-2, what="sibling %d dbeta" % k)
    dx = P8(N, Cin, *grid)
    ops.gn_backward(xp, dV, 8, stats, ones, None, None, False, dx, False, s2d=True)
    torch.cuda.synchronize()
    close(dx.to_ncdhw(), x.grad, tol=2e-2, what="summed dL/dx through the shared view")


CONAT_CASES = [
    # N, grid (Z, Y, X), channels of the four branches, Cout
    (2, (8, 48, 24), (32, 32, 64, 64), 128),
    (1, (16, 64, 160), (32, 32, 64, 64), 128),   # the BASELINE grid
    (3, (4, 44, 10), (16, 32), 48),              # two branches, Cout not a multiple of 32, coarsest z extent 2
]


@pytest.mark.parametrize("case", CONAT_CASES, ids=str)
def test_final_concat_conv_as_one_gemm(ctx, case):
    """cat(x0, trilinear-upsampled x1..x3) -> 1x1 conv + bias (backbones/hrnet3d.py:37-42) in one launch with the concat
    assembled in shared memory (csrc/conat.cu), against torch (F.interpolate align_corners=True + cat + conv3d) and against
    the per-branch route (1x1 convs + fuse_sum) it replaces."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, grid, chans, Cout = case
    xs = []
    for j, c in enumerate(chans):
        g = tuple(max(1, v >> j) for v in grid)
        xs.append(rnd(N, c, *g, seed=300 + j).cuda())
    w = rnd(Cout, sum(chans), 1, 1, 1, seed=310, scale=0.1).cuda()
    b = rnd(Cout, seed=311).cuda()
    cat = torch.cat([xs[0]] + [F.interpolate(x, size=grid, mode="trilinear", align_corners=True) for x in xs[1:]], 1)
    ref = F.conv3d(cat, w, b)
    ys = [P8.from_ncdhw(x) for x in xs]
    out = ops.conat_forward(ctx, ys, w, b, P8(N, Cout, *grid))
    assert out is not None, "shape not taken by rtp_conat_fwd"
    torch.cuda.synchronize()
    close(out.to_ncdhw(), ref, tol=1.5 * BF16_ULP, what="concat + 1x1 conv in one GEMM")
    # pads of the output stay zero (ring positions are computed but never stored)
    full = out.buf[out.offset:out.offset + out.N * out.n_stride].view(N, out.C8, grid[0], grid[2] + 2, grid[1] + 2, 8).float()
    assert float(full[:, :, :, 0].abs().sum()) == 0 and float(full[:, :, :, :, 0].abs().sum()) == 0
    assert float(full[:, :, :, -1].abs().sum()) == 0 and float(full[:, :, :, :, -1].abs().sum()) == 0
    # the route it replaces
    terms, c0 = [], 0
    for y in ys:
        t = P8(N, Cout, *y.grid)
        ops.conv_forward(ctx, y, w, 1, t, ci0=c0, ci_n=y.C)
        terms.append(t)
        c0 += y.C
    old = ops.fuse_sum(P8(N, Cout, *grid), [terms[0]], terms[1:], bias=b)
    torch.cuda.synchronize()
    e_new = (out.to_ncdhw() - ref).abs().max().item()
    e_old = (old.to_ncdhw() - ref).abs().max().item()
    print("max err vs torch: one-GEMM %.4g, per-branch convs + fuse_sum %.4g (ref max %.3g)" % (e_new, e_old, ref.abs().max().item()))
    assert e_new <= 1.5 * e_old + 1e-3


def test_final_concat_conv_falls_back_on_unsupported_shapes(ctx):
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    ys = [P8(1, 32, 8, 16, 24), P8(1, 32, 4, 8, 12)]   # rows of 18 padded positions: a tile would span > 4 rows
    w = torch.zeros(64, 64, 1, 1, 1, device="cuda")
    assert ops.conat_forward(ctx, ys, w, None, P8(1, 64, 8, 16, 24)) is None
    ys = [P8(1, 64, 8, 48, 24), P8(1, 128, 4, 24, 12), P8(1, 128, 2, 12, 6), P8(1, 64, 1, 6, 3)]  # feat64: weights exceed shared memory
    w = torch.zeros(256, 384, 1, 1, 1, device="cuda")
    assert ops.conat_forward(ctx, ys, w, None, P8(1, 256, 8, 48, 24)) is None


def test_fused_launch_variants_match_their_multi_launch_forms(ctx):
    """rtp_gn_stats == rtp_gn_sums + rtp_gn_finalize and rtp_conv_multi == one rtp_conv per parity class, bit for bit."""
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8, _stream
    x = to_p8(rnd(3, 64, 4, 10, 12, seed=90))
    a = ops.gn_stats(x, 8)
    sums = torch.empty((x.N, x.C, 2), dtype=torch.float32, device="cuda")
    b = torch.empty((x.N, 8, 2), dtype=torch.float32, device="cuda")
    lib.call("rtp_gn_sums", x.struct(), x.C, sums.data_ptr(), ops.gn_ws(x).data_ptr(), _stream())
    lib.call("rtp_gn_finalize", sums.data_ptr(), x.N, x.C, 8, x.voxels, 1e-5, b.data_ptr(), _stream())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # stride-2 dgrad through the gather kernel: 8 parity classes in one launch vs one launch each
    dy = to_p8(rnd(2, 32, 2, 5, 6, seed=91))
    w = rnd(32, 32, 3, 3, 3, seed=92, scale=0.1).cuda()
    old = ops.USE_S2D
    ops.USE_S2D = False
    try:
        d_multi = ops.conv_dgrad(ctx, dy, w, 2, P8(2, 32, 4, 10, 12))
    finally:
        ops.USE_S2D = old
    wp, KP, NP = ctx.get(w, 1)
    d_single = P8(2, 32, 4, 10, 12)
    for pz in range(2):
        for px in range(2):
            for py in range(2):
                rows = ((4 - pz + 1) // 2, (12 - px + 1) // 2, (10 - py + 1) // 2)
                ops.conv(dy, wp, KP, NP, d_single, ops.taps_dgrad_s2(pz, px, py), rows, IS=1, OS=2, off=(pz, px, py))
    torch.cuda.synchronize()
    assert torch.equal(d_multi.to_ncdhw(), d_single.to_ncdhw())


def test_head_loss_without_positives():
    """Frames whose skeleton falls outside the ROI: mask all zero, empty target heat-map (center_head.py:244-270 with
    num_pos = 0: hm_loss = -neg, regression loss 0 and no regression gradient)."""
    from oracle import hrpose_oracle as O
    from rtpose_b200.engine import Engine
    grid, N, ncls, R = (4, 8, 12), 2, 1, 45
    rs = np.random.RandomState(9)
    tgt = O.batch_targets([O.synth_pose(rs, grid) for _ in range(N)], grid, True)
    tgt = {k: v.clone() for k, v in tgt.items()}
    tgt["mask"].zero_()
    tgt["hm"].zero_()
    hm = bf(rnd(N, ncls, *grid, seed=40) * 0.5 - 2.0).requires_grad_(True)
    reg = rnd(N, R, *grid, seed=41).requires_grad_(True)
    L = O.head_loss({"hm": hm, "reg": reg}, tgt, 0.5, [1.0] * 45)
    L["loss"].backward()
    eng = Engine("hr_tiny_feat32_zyx_l4_in32", "top", {}, R, ncls, 0.5, [1.0] * 45)
    eng.begin()
    hp, rp = to_p8(hm.detach()), to_p8(reg.detach())
    out = eng.loss(hp, rp, tgt["hm"].cuda(), tgt["ind"].cuda(), tgt["mask"].cuda(), tgt["cat"].cuda(), tgt["anno_pose"].cuda())
    torch.cuda.synchronize()
    o = out.cpu()
    assert o[3].item() == 0.0 and L["num_positive"].item() == 0
    assert abs(o[0].item() - L["loss"].item()) <= 1e-4 * abs(L["loss"].item())
    assert o[2].item() == 0.0
    close(hp.grad.to_ncdhw(), hm.grad, tol=2 * BF16_ULP, what="d loss / d hm (no positives)")
    assert float(rp.grad.to_ncdhw().abs().max()) == 0.0


def test_full_size_linearity_properties(ctx):
    """Size-independent properties at the BASELINE shape (batch 16, 32 ch, 16x64x160), where the oracle is too slow to
    run: scaling an operand by 2 is exact in bf16 / fp32, so conv, dgrad and wgrad must scale bit-exactly; and the
    forward conv of a one-hot input reproduces the packed weights (checks tap / channel ordering at full resolution)."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cc, grid = 16, 32, (16, 64, 160)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, Cc, *grid, device="cuda", generator=g)
    w = (torch.randn(Cc, Cc, 3, 3, 3, device="cuda", generator=g) * 0.05)
    xp, x2p = P8.from_ncdhw(x), P8.from_ncdhw(2 * x)
    y1 = ops.conv_forward(ctx, xp, w, 1, P8(N, Cc, *grid)).to_ncdhw()
    y2 = ops.conv_forward(ctx, x2p, w, 1, P8(N, Cc, *grid)).to_ncdhw()
    assert torch.equal(y2, 2 * y1)
    d1 = ops.conv_dgrad(ctx, xp, w, 1, P8(N, Cc, *grid)).to_ncdhw()
    d2 = ops.conv_dgrad(ctx, x2p, w, 1, P8(N, Cc, *grid)).to_ncdhw()
    assert torch.equal(d2, 2 * d1)
    g1, g2 = torch.zeros_like(w), torch.zeros_like(w)
    ops.conv_wgrad(xp, xp, 3, 1, g1)
    ops.conv_wgrad(xp, x2p, 3, 1, g2)
    torch.cuda.synchronize()
    assert torch.equal(g2, 2 * g1)
    del y2, d1, d2, x2p
    # one-hot probe: x = delta at an interior voxel of sample 3, channel 5  ->  y[3, :, v - (k - 1)] = bf16(w[:, 5, k])
    z0, y0, x0 = 7, 31, 80
    xh = torch.zeros(N, Cc, *grid, device="cuda")
    xh[3, 5, z0, y0, x0] = 1.0
    yh = ops.conv_forward(ctx, P8.from_ncdhw(xh), w, 1, P8(N, Cc, *grid)).to_ncdhw()
    patch = yh[3, :, z0 - 1:z0 + 2, y0 - 1:y0 + 2, x0 - 1:x0 + 2]
    assert torch.equal(patch, bf(w[:, 5].flip(1, 2, 3).cpu()).cuda())
    assert float(yh.abs().sum()) == float(patch.abs().sum())


def test_wgrad_s2d_wide_output_as_channel_slices():
    """Stride-2 conv 32 -> 64 through the space-to-depth view: the plane-streaming weight gradient takes dY in 32-channel
    slices (dW rows [32 h, 32 h + 32)); against the gather kernel on the plain tensor and against itself with the slicing
    switched off (generic kernel over the view)."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cin, Cout, grid, og = 4, 32, 64, (8, 32, 48), (4, 16, 24)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = P8.from_ncdhw(torch.randn(N, Cin, *grid, device="cuda", generator=g))
    dy = P8.from_ncdhw(torch.randn(N, Cout, *og, device="cuda", generator=g))
    gamma, beta = torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
    st = ops.gn_stats(x, 8)
    xn = ops.gn_apply(x, 8, st, gamma, beta, P8(N, Cin, *grid))
    xs = ops.gn_apply_s2d(x, 8, st, gamma, beta, P8(N, 8 * Cin, *og))
    gw_s, gw_v, gw_g = (torch.zeros(Cout, Cin, 3, 3, 3, device="cuda") for _ in range(3))
    lib_calls = []
    real = ops.lib.call
    ops.lib.call = lambda name, *a: (lib_calls.append(name), real(name, *a))[1]
    try:
        ops.conv_wgrad_s2d(xs, dy, Cin, gw_s)
    finally:
        ops.lib.call = real
    assert lib_calls.count("rtp_wgrad_s2d") == 2, lib_calls
    old = ops.USE_WGRAD_S2D
    ops.USE_WGRAD_S2D = False
    try:
        ops.conv_wgrad_s2d(xs, dy, Cin, gw_v)
    finally:
        ops.USE_WGRAD_S2D = old
    ops.conv_wgrad(xn, dy, 3, 2, gw_g)
    torch.cuda.synchronize()
    close(gw_s, gw_g, tol=1e-3, what="sliced s2d wgrad vs gather kernel")
    close(gw_s, gw_v, tol=1e-4, what="sliced s2d wgrad vs generic kernel over the view")
    ops.conv_wgrad_s2d(xs, dy, Cin, gw_s, accumulate=True)
    torch.cuda.synchronize()
    close(gw_s, 2 * gw_g, tol=1e-3, what="sliced s2d wgrad, accumulate")


def test_full_size_stride2_s2d_agrees_with_gather_kernels(ctx):
    """At the BASELINE shape the stride-2 exchange conv is computed twice with independent kernels — through the
    space-to-depth view (plane-streaming conv, masked taps, paired dgrad launches, view wgrad) and with the gather
    kernels on the plain tensor — and the results must agree to bf16 / fp32-accumulation round-off."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    N, Cc, grid, og = 16, 32, (16, 64, 160), (8, 32, 80)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = P8.from_ncdhw(torch.randn(N, Cc, *grid, device="cuda", generator=g))
    dy = P8.from_ncdhw(torch.randn(N, Cc, *og, device="cuda", generator=g))
    w = torch.randn(Cc, Cc, 3, 3, 3, device="cuda", generator=g) * 0.05
    gamma, beta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    st = ops.gn_stats(x, 8)
    assert ops.s2d_eligible(x, w)
    xn = ops.gn_apply(x, 8, st, gamma, beta, P8(N, Cc, *grid))
    xs = ops.gn_apply_s2d(x, 8, st, gamma, beta, P8(N, 8 * Cc, *og))
    we = ops.s2d_expand(w)
    y_s = ops.conv_forward(ctx, xs, we, 1, P8(N, Cc, *og), key="fs2", version=0,
                           tap_mask=[ops.s2d_tap_mask(p, False) for p in range(8)]).to_ncdhw()
    y_g = ops.conv_forward(ctx, xn, w, 2, P8(N, Cc, *og)).to_ncdhw()
    close(y_s, y_g, what="s2 forward: s2d vs gather")
    gw_s, gw_g = torch.zeros_like(w), torch.zeros_like(w)
    ops.conv_wgrad_s2d(xs, dy, Cc, gw_s)
    ops.conv_wgrad(xn, dy, 3, 2, gw_g)
    torch.cuda.synchronize()
    close(gw_s, gw_g, tol=1e-3, what="s2 wgrad: view vs plain tensor")
    dxs = ops.conv_dgrad(ctx, dy, we, 1, P8(N, 8 * Cc, *og), key="fs2", version=0, s2d_cin=Cc)
    old = ops.USE_S2D
    ops.USE_S2D = False
    try:
        dxn = ops.conv_dgrad(ctx, dy, w, 2, P8(N, Cc, *grid))
    finally:
        ops.USE_S2D = old
    # compare through GroupNorm backward (reads the view / the plain gradient)
    dg1, db1, dg2, db2 = (torch.zeros(Cc, device="cuda") for _ in range(4))
    dx1, dx2 = P8(N, Cc, *grid), P8(N, Cc, *grid)
    ops.gn_backward(x, dxs, 8, st, gamma, dg1, db1, False, dx1, False, s2d=True)
    ops.gn_backward(x, dxn, 8, st, gamma, dg2, db2, False, dx2, False)
    torch.cuda.synchronize()
    close(dx1.to_ncdhw(), dx2.to_ncdhw(), tol=2 * BF16_ULP, what="s2 dgrad + GN backward: s2d vs gather")
    close(dg1, dg2, tol=1e-3, what="dgamma")
    close(db1, db2, tol=1e-3, what="dbeta")


@pytest.mark.parametrize("case", [(2, 32, 128, (3, 20, 18)), (1, 64, 128, (2, 14, 30)), (2, 32, 32, (4, 16, 14)), (1, 40, 24, (2, 14, 16))],
                         ids=str)
def test_pointwise_wgrad_streaming_kernel(case):
    """rtp_wgrad_pw (1x1x1 weight gradient as a streaming GEMM over the padded positions) against torch and against the
    gather kernel, with the final-conv style input-channel offset (ci0) and accumulation."""
    from rtpose_b200 import lib, ops
    N, Cin, Cout, grid = case
    x = rnd(N, Cin, *grid, seed=21)
    dy = rnd(N, Cout, *grid, seed=22)
    w = torch.zeros(Cout, Cin, 1, 1, 1, requires_grad=True)
    F.conv3d(x, w).backward(dy)
    xp, dyp = to_p8(x), to_p8(dy)
    assert lib.load().rtp_wgrad_pw_supported(Cin, Cout, *[grid[0], grid[2], grid[1]]) == 1
    lib.call_counts.clear()
    big = torch.full((Cout, Cin + 24, 1, 1, 1), 3.0, device="cuda")
    ops.conv_wgrad(xp, dyp, 1, 1, big, ci0=16)
    ops.conv_wgrad(xp, dyp, 1, 1, big, ci0=16, accumulate=True)
    torch.cuda.synchronize()
    assert lib.call_counts.get("rtp_wgrad_pw", 0) == 2
    close(big[:, 16:16 + Cin], 2 * w.grad, tol=2e-3, what="streaming 1x1 wgrad (ci0, accumulate)")
    assert float((big[:, :16] - 3.0).abs().max()) == 0 and float((big[:, 16 + Cin:] - 3.0).abs().max()) == 0
    old = ops.USE_WGRAD_PW
    ops.USE_WGRAD_PW = False
    try:
        ref = torch.zeros((Cout, Cin, 1, 1, 1), device="cuda")
        ops.conv_wgrad(xp, dyp, 1, 1, ref)
    finally:
        ops.USE_WGRAD_PW = old
    got = torch.zeros((Cout, Cin, 1, 1, 1), device="cuda")
    ops.conv_wgrad(xp, dyp, 1, 1, got)
    torch.cuda.synchronize()
    close(got, ref, tol=2e-5, what="streaming vs gather 1x1 wgrad (same bf16 operands)")
    # bias gradient as a ones channel of the same GEMM (rtp_wgrad_pw_bias): dW unchanged, db = sum over positions of dy
    lib.call_counts.clear()
    got_b = torch.zeros((Cout, Cin, 1, 1, 1), device="cuda")
    db = torch.full((Cout,), 5.0, device="cuda")
    ops.conv_wgrad(xp, dyp, 1, 1, got_b, bias_grad=(db, False))
    torch.cuda.synchronize()
    assert lib.call_counts.get("rtp_wgrad_pw_bias", 0) == 1 and lib.call_counts.get("rtp_channel_sum", 0) == 0
    assert torch.equal(got_b, got)
    want_db = dy.sum((0, 2, 3, 4))
    close(db, want_db, tol=1e-4, what="bias gradient from the ones channel")
    ops.conv_wgrad(xp, dyp, 1, 1, got_b, accumulate=True, bias_grad=(db, True))
    torch.cuda.synchronize()
    close(db, 2 * want_db, tol=1e-4, what="bias gradient, accumulated")
    close(got_b, 2 * got, tol=1e-6, what="dW accumulated beside the bias column")
