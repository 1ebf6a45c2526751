"""Host-side check of the GroupNorm-backward algebra the device kernels implement (csrc/norm.cu gn_bwd_partial_kernel /
gn_bwd_apply_kernel, and the conv epilogue's fused reductions): per (sample, channel) reductions
red0 = sum dy, red1 = sum dy * xhat; per group s1 = sum_c gamma_c red0, s2 = sum_c gamma_c red1, m = voxels * channels/group;
dx = rstd * (gamma * dy - s1/m - xhat * s2/m); dbeta = sum_n red0, dgamma = sum_n red1 — against autograd of
torch.nn.functional.group_norm (the reference's nn.GroupNorm, hr_util/common.py:57) in float64, with and without the ReLU
mask of the 'gcr' blocks."""
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("C,G,relu", [(32, 8, False), (64, 8, True), (4, 1, False)])
def test_group_norm_backward_algebra(C, G, relu):
    g = torch.Generator().manual_seed(C + G)
    x = torch.randn(3, C, 4, 5, 6, generator=g, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(C, generator=g, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(C, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(3, C, 4, 5, 6, generator=g, dtype=torch.float64)
    y = F.group_norm(x, G, gamma, beta, eps=1e-5)
    y.backward(dy)
    # the kernels' formulation
    xd = x.detach()
    N, cpg, V = 3, C // G, 4 * 5 * 6
    xg = xd.view(N, G, cpg * V)
    mean, var = xg.mean(dim=2), xg.var(dim=2, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    xhat = ((xg - mean[..., None]) * rstd[..., None]).view_as(xd)
    red0, red1 = dy.sum(dim=(2, 3, 4)), (dy * xhat).sum(dim=(2, 3, 4))          # [N, C]
    s1 = (gamma.detach() * red0).view(N, G, cpg).sum(dim=2)
    s2 = (gamma.detach() * red1).view(N, G, cpg).sum(dim=2)
    m = float(V * cpg)
    ex = lambda t: t.repeat_interleave(cpg, dim=1)[:, :, None, None, None]      # group value -> its channels
    dx = ex(rstd) * (gamma.detach()[None, :, None, None, None] * dy - ex(s1) / m - xhat * ex(s2) / m)
    torch.testing.assert_close(dx, x.grad, rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(red0.sum(0), beta.grad, rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(red1.sum(0), gamma.grad, rtol=1e-10, atol=1e-10)
    if relu:  # x is itself a ReLU output whose gradient is kept pre-mask: dx *= (x > 0)
        x2 = torch.randn(3, C, 4, 5, 6, generator=g, dtype=torch.float64, requires_grad=True)
        F.group_norm(F.relu(x2), G, gamma.detach(), beta.detach(), eps=1e-5).backward(dy)
        xr = F.relu(x2.detach())
        xg = xr.view(N, G, cpg * V)
        mean, rstd = xg.mean(dim=2), (xg.var(dim=2, unbiased=False) + 1e-5).rsqrt()
        xhat = ((xg - mean[..., None]) * rstd[..., None]).view_as(xr)
        r0, r1 = dy.sum(dim=(2, 3, 4)), (dy * xhat).sum(dim=(2, 3, 4))
        t1 = (gamma.detach() * r0).view(N, G, cpg).sum(dim=2)
        t2 = (gamma.detach() * r1).view(N, G, cpg).sum(dim=2)
        dxr = ex(rstd) * (gamma.detach()[None, :, None, None, None] * dy - ex(t1) / m - xhat * ex(t2) / m) * (xr > 0)
        torch.testing.assert_close(dxr, x2.grad, rtol=1e-10, atol=1e-10)
