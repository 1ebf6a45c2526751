"""DCN v1 operator parity: rtp_dcn_* against torchvision.ops.deform_conv2d (CPU, mask=None), which
implements the same DCNv1 arithmetic and offset-channel order [dg][kh*kw][dy,dx] as the reference kernels
(det3d/ops/dcn/src/deform_conv_cuda_kernel.cu:84-115,190-243).  fp32 both sides: rtol 1e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu
tv = pytest.importorskip("torchvision.ops")

CASES = [(2, 8, 7, 9, 8, 3, 1, 1, 1, 4), (1, 16, 12, 10, 24, 3, 1, 1, 1, 4), (2, 8, 9, 11, 16, 3, 2, 1, 1, 2),
         (1, 32, 16, 20, 32, 3, 1, 1, 1, 4), (1, 8, 8, 8, 8, 1, 1, 0, 1, 1), (1, 8, 10, 9, 8, 3, 1, 2, 2, 1)]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_deform_conv_matches_torchvision(case):
    from rtpose_b200.dcn import DeformConv
    N, Cc, H, W, Cout, k, stride, pad, dil, dg = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cc, H, W, generator=g, requires_grad=True)
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    off = (torch.randn(N, dg * 2 * k * k, Ho, Wo, generator=g) * 1.5).requires_grad_(True)
    m = DeformConv(Cc, Cout, k, stride=stride, padding=pad, dilation=dil, deformable_groups=dg)
    w = m.weight.detach().clone().requires_grad_(True)
    ref = tv.deform_conv2d(x, off, w, None, stride=stride, padding=pad, dilation=dil)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    m = m.cuda()
    xc, oc = x.detach().cuda().requires_grad_(True), off.detach().cuda().requires_grad_(True)
    out = m(xc, oc)
    out.backward(gy.cuda())
    torch.cuda.synchronize()
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(xc.grad.cpu(), x.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(oc.grad.cpu(), off.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(m.weight.grad.cpu(), w.grad, rtol=1e-4, atol=1e-3)


def test_deform_conv_rejects_cpu_and_bad_rank():
    from rtpose_b200.dcn import deform_conv
    with pytest.raises(NotImplementedError):
        deform_conv(torch.zeros(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.zeros(4, 4, 3, 3), 1, 1, 1, 1, 1)
    with pytest.raises(ValueError):
        deform_conv(torch.zeros(1, 4, 2, 5, 5).cuda(), torch.zeros(1, 18, 5, 5).cuda(), torch.zeros(4, 4, 3, 3).cuda(), 1, 1, 1, 1, 1)
