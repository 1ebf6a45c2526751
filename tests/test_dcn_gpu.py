"""DCN v1 operator parity: rtp_dcn_* against torchvision.ops.deform_conv2d (CPU, mask=None), which
implements the same DCNv1 arithmetic and offset-channel order [dg][kh*kw][dy,dx] as the reference kernels
(det3d/ops/dcn/src/deform_conv_cuda_kernel.cu:84-115,190-243).  fp32 both sides: rtol 1e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu
tv = pytest.importorskip("torchvision.ops")


@pytest.fixture(autouse=True)
def _fp32_kernels_by_default(monkeypatch):
    """The operator-parity tests below pin the fp32 CUDA-core kernels (rtol 1e-4 against torchvision); the tensor-core path —
    the library default — is selected explicitly by the tests that check it (bf16 tolerances)."""
    from rtpose_b200 import dcn
    monkeypatch.setattr(dcn, "TENSOR_CORE", False)
    monkeypatch.setattr(dcn, "TENSOR_CORE_BACKWARD", False)

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


MCASES = [(2, 8, 7, 9, 8, 3, 1, 1, 1, 4, True), (1, 16, 12, 10, 24, 3, 1, 1, 1, 4, False), (2, 8, 9, 11, 16, 3, 2, 1, 1, 2, True),
          (1, 8, 8, 8, 8, 1, 1, 0, 1, 1, True), (1, 8, 10, 9, 8, 3, 1, 2, 2, 1, False), (1, 32, 16, 40, 32, 3, 1, 1, 1, 4, True)]


@pytest.mark.parametrize("case", MCASES, ids=[str(c) for c in MCASES])
def test_modulated_deform_conv_matches_torchvision(case):
    """DCN v2 (det3d/ops/dcn/deform_conv.py:115-186, kernels deform_conv_cuda_kernel.cu:571-767): rtp_mdcn_* against
    torchvision.ops.deform_conv2d with a mask (same arithmetic, mask channel order [dg][kh*kw])."""
    from rtpose_b200.dcn import ModulatedDeformConv
    N, Cc, H, W, Cout, k, stride, pad, dil, dg, with_bias = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cc, H, W, generator=g, requires_grad=True)
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    off = (torch.randn(N, dg * 2 * k * k, Ho, Wo, generator=g) * 1.5).requires_grad_(True)
    mask = torch.sigmoid(torch.randn(N, dg * k * k, Ho, Wo, generator=g)).requires_grad_(True)
    m = ModulatedDeformConv(Cc, Cout, k, stride=stride, padding=pad, dilation=dil, deformable_groups=dg, bias=with_bias)
    if with_bias:
        m.bias.data.copy_(torch.randn(Cout, generator=g))
    w = m.weight.detach().clone().requires_grad_(True)
    b = m.bias.detach().clone().requires_grad_(True) if with_bias else None
    ref = tv.deform_conv2d(x, off, w, b, stride=stride, padding=pad, dilation=dil, mask=mask)
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    m = m.cuda()
    xc, oc, mc = (t.detach().cuda().requires_grad_(True) for t in (x, off, mask))
    out = m(xc, oc, mc)
    out.backward(gy.cuda())
    torch.cuda.synchronize()
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(xc.grad.cpu(), x.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(oc.grad.cpu(), off.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(mc.grad.cpu(), mask.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(m.weight.grad.cpu(), w.grad, rtol=1e-4, atol=1e-3)
    if with_bias:
        torch.testing.assert_close(m.bias.grad.cpu(), b.grad, rtol=1e-4, atol=1e-3)


def test_far_outside_and_nan_offsets_never_become_addresses():
    """Sampling positions far outside the image (or NaN) are 'invalid' (deform_conv_cuda_kernel.cu:229): they contribute
    0 and must not be dereferenced.  x sits at the very end of its allocation so a stray read would fault or pick up the
    NaN guard that follows it."""
    from rtpose_b200.dcn import deform_conv
    g = torch.Generator().manual_seed(2)
    buf = torch.full((2 * 8 * 6 * 7 + 4096,), float("nan")).cuda()
    x = buf[:2 * 8 * 6 * 7].view(2, 8, 6, 7)
    x.copy_(torch.randn(2, 8, 6, 7, generator=g))
    w = torch.randn(4, 8, 3, 3, generator=g).cuda()
    off = torch.zeros(2, 18, 6, 7)
    off[:, 0::2] = torch.tensor([1e4, -1e4, 6.0, 40.0, -7.5, 5.5, 1e9, float("nan"), 0.25])[None, :, None, None]
    off[:, 1::2] = torch.tensor([0.0, 3.0, 1e4, 0.5, -1e9, 6.5, 7.0, 0.0, float("nan")])[None, :, None, None]
    out = deform_conv(x, off.cuda(), w, 1, 1, 1, 1, 1)
    ref = tv.deform_conv2d(x.cpu(), torch.nan_to_num(off, nan=1e9), w.cpu(), None, padding=1)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-4)


def test_mask_of_ones_is_v1_bitwise():
    from rtpose_b200.dcn import deform_conv, modulated_deform_conv
    g = torch.Generator().manual_seed(9)
    x, w = torch.randn(2, 16, 10, 12, generator=g).cuda(), torch.randn(8, 16, 3, 3, generator=g).cuda()
    off = (torch.randn(2, 4 * 18, 10, 12, generator=g) * 2).cuda()
    a = deform_conv(x, off, w, 1, 1, 1, 1, 4)
    b = modulated_deform_conv(x, off, torch.ones(2, 4 * 9, 10, 12).cuda(), w, None, 1, 1, 1, 1, 4)
    assert torch.equal(a, b)


@pytest.mark.parametrize("modulated", [False, True])
def test_pack_modules(modulated):
    """*Pack modules (deform_conv.py:258-323, :382-446): with the zero-initialised predictor the op is a plain conv
    (x 0.5 for v2: sigmoid(0)); with a trained predictor it equals predictor -> torchvision deform_conv2d."""
    import torch.nn.functional as F
    from rtpose_b200.dcn import DeformConvPack, ModulatedDeformConvPack
    torch.manual_seed(1)
    m = (ModulatedDeformConvPack(8, 16, 3, padding=1, deformable_groups=2, bias=True) if modulated
         else DeformConvPack(8, 16, 3, padding=1, deformable_groups=2)).cuda()
    x = torch.randn(2, 8, 9, 11).cuda()
    plain = F.conv2d(x.cpu(), m.weight.detach().cpu(), None, padding=1)
    torch.testing.assert_close(m(x).detach().cpu(), plain * (0.5 if modulated else 1.0), rtol=1e-4, atol=1e-4)
    m.conv_offset.weight.data.normal_(0, 0.1)
    m.conv_offset.bias.data.normal_(0, 0.5)
    pred = F.conv2d(x.cpu(), m.conv_offset.weight.detach().cpu(), m.conv_offset.bias.detach().cpu(), padding=1)
    if modulated:
        o1, o2, mk = torch.chunk(pred, 3, dim=1)
        ref = tv.deform_conv2d(x.cpu(), torch.cat((o1, o2), 1), m.weight.detach().cpu(), m.bias.detach().cpu(), padding=1,
                               mask=torch.sigmoid(mk))
    else:
        ref = tv.deform_conv2d(x.cpu(), pred, m.weight.detach().cpu(), None, padding=1)
    out = m(x)
    out.sum().backward()
    assert m.conv_offset.weight.grad is not None and m.weight.grad is not None
    torch.testing.assert_close(out.detach().cpu(), ref, rtol=2e-3, atol=2e-3)  # the predictor conv runs in TF32/fp32 on the GPU


def test_modulated_rejects_cpu_and_bad_shapes():
    from rtpose_b200.dcn import modulated_deform_conv
    with pytest.raises(NotImplementedError):
        modulated_deform_conv(torch.zeros(1, 4, 5, 5), torch.zeros(1, 18, 5, 5), torch.zeros(1, 9, 5, 5), torch.zeros(4, 4, 3, 3), None, 1, 1)
    z = lambda *s: torch.zeros(*s).cuda()
    with pytest.raises(ValueError, match="mask shape"):
        modulated_deform_conv(z(1, 4, 5, 5), z(1, 18, 5, 5), z(1, 8, 5, 5), z(4, 4, 3, 3), None, 1, 1)
    with pytest.raises(ValueError, match="offset shape"):
        modulated_deform_conv(z(1, 4, 5, 5), z(1, 18, 4, 5), z(1, 9, 5, 5), z(4, 4, 3, 3), None, 1, 1)


TC_CASES = [(2, 32, 12, 20, 32, 3, 1, 1, 1, 4, False), (3, 64, 16, 40, 48, 3, 1, 1, 1, 4, True), (1, 128, 64, 160, 128, 3, 1, 1, 1, 4, False),
            (2, 16, 9, 11, 24, 3, 2, 1, 1, 2, True), (1, 32, 10, 9, 16, 1, 1, 0, 1, 1, True)]


@pytest.mark.parametrize("case", TC_CASES, ids=[str(c) for c in TC_CASES])
def test_tensor_core_forward(case, monkeypatch):
    """RTP_DCN_TC path (rtp_dcn_sample_p8 + rtp_conv on tcgen05): bf16 operands, fp32 accumulation -> the conv tolerance of
    SURVEY.md §8c (1), |d| <= 2^-7 * max|ref|, against torchvision's fp32 op; v1 and v2 (mask, bias); the batch is also
    pushed through in chunks (TC_SAMPLE_BYTES) and must give the same bits."""
    from rtpose_b200 import dcn
    N, Cc, H, W, Cout, k, stride, pad, dil, dg, modulated = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cc, H, W, generator=g)
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    off = torch.randn(N, dg * 2 * k * k, Ho, Wo, generator=g) * 1.5
    w = torch.randn(Cout, Cc, k, k, generator=g) / (Cc * k * k) ** 0.5
    mask = torch.sigmoid(torch.randn(N, dg * k * k, Ho, Wo, generator=g)) if modulated else None
    bias = torch.randn(Cout, generator=g) if modulated else None
    ref = tv.deform_conv2d(x, off, w, bias, stride=stride, padding=pad, dilation=dil, mask=mask)
    monkeypatch.setattr(dcn, "TENSOR_CORE", True)

    def run():
        if modulated:
            return dcn.modulated_deform_conv(x.cuda(), off.cuda(), mask.cuda(), w.cuda(), bias.cuda(), stride, pad, dil, 1, dg)
        return dcn.deform_conv(x.cuda(), off.cuda(), w.cuda(), stride, pad, dil, 1, dg)
    out = run()
    torch.cuda.synchronize()
    err, lim = (out.cpu() - ref).abs().max().item(), 2.0 ** -7 * ref.abs().max().item()
    assert err <= lim, "max abs err %.4g > %.4g" % (err, lim)
    if N > 1:
        per = -(-Cc // 8) * k * k * (Wo + 2) * (Ho + 2) * 16
        monkeypatch.setattr(dcn, "TC_SAMPLE_BYTES", per)  # one sample per chunk
        assert torch.equal(run(), out)
    # gradients still flow (fp32 kernels) when the forward ran on the tensor cores
    xg = x.cuda().requires_grad_(True)
    y = dcn.modulated_deform_conv(xg, off.cuda(), mask.cuda(), w.cuda(), bias.cuda(), stride, pad, dil, 1, dg) if modulated else \
        dcn.deform_conv(xg, off.cuda(), w.cuda(), stride, pad, dil, 1, dg)
    y.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()


@pytest.mark.parametrize("case", TC_CASES[:4], ids=[str(c) for c in TC_CASES[:4]])
def test_tensor_core_backward(case, monkeypatch):
    """RTP_DCN_TC_BWD path: weight gradient by rtp_wgrad over the sample volume, sample gradient by single-tap rtp_conv
    launches, scatter by rtp_dcn_col2im_p8 — against torchvision's fp32 autograd with the conv tolerance (bf16 operands,
    fp32 accumulation): |d| <= 2^-7 * max|ref| per gradient tensor (2^-6 for the offset gradient, a product of two
    rounded factors).  Chunked execution must give the same input-side gradients bit for bit."""
    from rtpose_b200 import dcn
    N, Cc, H, W, Cout, k, stride, pad, dil, dg, modulated = case
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, Cc, H, W, generator=g, requires_grad=True)
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    off = (torch.randn(N, dg * 2 * k * k, Ho, Wo, generator=g) * 1.5).requires_grad_(True)
    w = (torch.randn(Cout, Cc, k, k, generator=g) / (Cc * k * k) ** 0.5).requires_grad_(True)
    mask = torch.sigmoid(torch.randn(N, dg * k * k, Ho, Wo, generator=g)).requires_grad_(True) if modulated else None
    bias = torch.randn(Cout, generator=g).requires_grad_(True) if modulated else None
    gy = torch.randn(N, Cout, Ho, Wo, generator=g)
    tv.deform_conv2d(x, off, w, bias, stride=stride, padding=pad, dilation=dil, mask=mask).backward(gy)
    monkeypatch.setattr(dcn, "TENSOR_CORE", True)
    monkeypatch.setattr(dcn, "TENSOR_CORE_BACKWARD", True)

    def run():
        leaves = [t.detach().cuda().requires_grad_(True) if t is not None else None for t in (x, off, mask, w, bias)]
        xc, oc, mc, wc, bc = leaves
        y = dcn.modulated_deform_conv(xc, oc, mc, wc, bc, stride, pad, dil, 1, dg) if modulated else \
            dcn.deform_conv(xc, oc, wc, stride, pad, dil, 1, dg)
        y.backward(gy.cuda())
        torch.cuda.synchronize()
        return [t.grad.cpu() if t is not None else None for t in leaves]
    got = run()
    for name, a, r, tol in zip(("dx", "doffset", "dmask", "dw", "dbias"), got, (x, off, mask, w, bias),
                               (2.0 ** -7, 2.0 ** -6, 2.0 ** -7, 2.0 ** -7, 1e-5)):
        if r is None:
            continue
        err, lim = (a - r.grad).abs().max().item(), tol * r.grad.abs().max().item()
        assert err <= lim, "%s: max abs err %.4g > %.4g" % (name, err, lim)
    if N > 1:
        per = -(-Cc // 8) * k * k * (Wo + 2) * (Ho + 2) * 16
        monkeypatch.setattr(dcn, "TC_SAMPLE_BYTES", per)
        again = run()
        assert torch.equal(again[1], got[1]) and (mask is None or torch.equal(again[2], got[2]))  # per-sample quantities
        torch.testing.assert_close(again[3], got[3], rtol=1e-4, atol=1e-5)  # dw: split-K order differs across chunkings


@pytest.mark.parametrize("tc", [False, True])
def test_feature_adaption(tc, monkeypatch):
    """FeatureAdaption (center_head.py:24-62) = 1x1 offset predictor -> DCN v1 (dg 4) -> ReLU, against the same pipeline
    built from torch / torchvision ops with the module's own parameters; fp32 kernels and the tensor-core path."""
    import torch.nn.functional as F
    from rtpose_b200 import dcn
    monkeypatch.setattr(dcn, "TENSOR_CORE", tc)
    monkeypatch.setattr(dcn, "TENSOR_CORE_BACKWARD", tc)
    torch.manual_seed(3)
    m = dcn.FeatureAdaption(32, 32, kernel_size=3, deformable_groups=4)
    assert sorted(m.state_dict()) == ["conv_adaption.weight", "conv_offset.bias", "conv_offset.weight"]
    assert float(m.conv_offset.weight.abs().sum()) == 0.0 and m.conv_offset.weight.shape == (72, 32, 1, 1)
    m.conv_offset.weight.data.normal_(0, 0.2)  # a trained predictor
    x = torch.randn(2, 32, 12, 20)
    off = F.conv2d(x, m.conv_offset.weight, m.conv_offset.bias)
    ref = F.relu(tv.deform_conv2d(x, off, m.conv_adaption.weight, None, padding=1))
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    out = m(xc)
    out.sum().backward()
    torch.cuda.synchronize()
    tol = 2.0 ** -7 * ref.abs().max().item() if tc else 2e-3
    assert (out.detach().cpu() - ref).abs().max().item() <= tol
    assert xc.grad is not None and m.conv_offset.weight.grad is not None and m.conv_adaption.weight.grad is not None


def test_tensor_core_path_is_the_default():
    import importlib
    import os
    from rtpose_b200 import dcn
    if os.environ.get("RTP_DCN_TC") or os.environ.get("RTP_DCN_TC_BWD"):
        pytest.skip("explicitly configured")
    fresh = importlib.reload(dcn)
    assert fresh.TENSOR_CORE and fresh.TENSOR_CORE_BACKWARD


def test_dcn_head_fold_z_matches_torch_composition(monkeypatch):
    """CenterHead(dcn_head='fold_z') — the 3-D-compatible definition of the reference's DCNSepHead (2-D ops on the z-folded
    batch, center_head.py:24-62,111-163) — against the same composition written with torch / torchvision ops in fp32:
    FeatureAdaption x2 (1x1 offsets, deform_conv2d dg=4, ReLU), cls head (conv2d 3x3, GroupNorm(8), ReLU, conv2d 3x3),
    SepHead (conv3d 3x3x3, ReLU, conv3d).  Tensor-core path (bf16 operands): conv tolerances."""
    import torch.nn.functional as F
    from rtpose_b200 import dcn
    from rtpose_b200 import det3d_compat as D
    monkeypatch.setattr(dcn, "TENSOR_CORE", True)
    monkeypatch.setattr(dcn, "TENSOR_CORE_BACKWARD", True)
    B, Cc, Z, Y, X = 2, 32, 3, 12, 20
    names = ["Pelvis"]
    torch.manual_seed(0)
    head = D.CenterHead(in_channels=Cc, tasks=[dict(num_class=1, class_names=names)], dataset="cruw_pose", weight=0.5,
                        code_weights=[1.0] * 45, common_heads={"reg": (45, 2)}, share_conv_channel=Cc, dcn_head="fold_z").cuda()
    with pytest.raises(TypeError):
        D.CenterHead(in_channels=Cc, tasks=[dict(num_class=1, class_names=names)], common_heads={"reg": (45, 2)},
                     share_conv_channel=Cc, dcn_head=True)
    t = head.tasks[0]
    with torch.no_grad():  # non-trivial offsets (the reference initialises the offset weights to zero)
        for fa in (t.feature_adapt_cls, t.feature_adapt_reg):
            fa.conv_offset.weight.normal_(0, 0.05)
            fa.conv_offset.bias.uniform_(-1.0, 1.0)
    g = torch.Generator().manual_seed(5)
    bf = lambda v: v.to(torch.bfloat16).float()
    x = bf(torch.randn(B, Cc, Z, Y, X, generator=g)).cuda().requires_grad_(True)
    preds, same = head(x)
    hm, reg = preds[0]["hm"], preds[0]["reg"]
    assert hm.shape == (B, 1, Z, Y, X) and reg.shape == (B, 45, Z, Y, X) and same is x
    gh, gr = torch.randn(hm.shape, generator=g).cuda(), torch.randn(reg.shape, generator=g).cuda()
    (hm * gh).sum().add((reg * gr).sum()).backward()
    got = {k: p.grad.clone() for k, p in head.named_parameters()}
    gx = x.grad.clone()
    # ---- torch / torchvision composition, fp32
    p = {k: v.detach().clone().requires_grad_(True) for k, v in head.named_parameters()}
    xr = x.detach().clone().requires_grad_(True)

    def adapt(prefix, x5):
        x2 = x5.permute(0, 2, 1, 3, 4).reshape(B * Z, Cc, Y, X)
        off = F.conv2d(x2, p[prefix + ".conv_offset.weight"], p[prefix + ".conv_offset.bias"])
        y2 = F.relu(tv.deform_conv2d(x2, off, p[prefix + ".conv_adaption.weight"], None, padding=1))
        return y2, y2.reshape(B, Z, Cc, Y, X).permute(0, 2, 1, 3, 4)
    c2, _ = adapt("tasks.0.feature_adapt_cls", xr)
    _, r5 = adapt("tasks.0.feature_adapt_reg", xr)
    h = F.conv2d(c2, p["tasks.0.cls_head.0.weight"], p["tasks.0.cls_head.0.bias"], padding=1)
    # GroupNorm over the whole (z, y, x) volume of a sample, like every norm of the path: un-fold, normalise, fold back
    h5 = h.reshape(B, Z, -1, Y, X).permute(0, 2, 1, 3, 4)
    h5 = F.relu(F.group_norm(h5, 8, p["tasks.0.cls_head.1.weight"], p["tasks.0.cls_head.1.bias"], 1e-5))
    h = h5.permute(0, 2, 1, 3, 4).reshape(B * Z, -1, Y, X)
    hm_r = F.conv2d(h, p["tasks.0.cls_head.3.weight"], p["tasks.0.cls_head.3.bias"], padding=1)
    hm_r = hm_r.reshape(B, Z, 1, Y, X).permute(0, 2, 1, 3, 4)
    tr = F.relu(F.conv3d(r5, p["tasks.0.task_head.reg.0.weight"], p["tasks.0.task_head.reg.0.bias"], padding=1))
    reg_r = F.conv3d(tr, p["tasks.0.task_head.reg.2.weight"], p["tasks.0.task_head.reg.2.bias"], padding=1)
    (hm_r * gh).sum().add((reg_r * gr).sum()).backward()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
    print("hm rel err %.4g, reg rel err %.4g, dx rel err %.4g" % (rel(hm, hm_r), rel(reg, reg_r), rel(gx, xr.grad)))
    def cosine(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float(a @ b / (a.norm() * b.norm() + 1e-30))
    cosines = {k: cosine(got[k], p[k].grad) for k in p}
    print("dx cosine %.5f; parameter-gradient cosines: %s" % (cosine(gx, xr.grad), {k[8:]: round(v, 4) for k, v in cosines.items()}))
    assert rel(hm, hm_r) <= 2 ** -5 and rel(reg, reg_r) <= 2 ** -5
    assert cosine(gx, xr.grad) >= 0.99
    for k in p:
        assert got[k] is not None, k
        assert cosines[k] >= 0.99, (k, cosines[k])
    assert sorted(k for k in dict(head.named_parameters())) == sorted(
        ["tasks.0.feature_adapt_cls.conv_offset.weight", "tasks.0.feature_adapt_cls.conv_offset.bias",
         "tasks.0.feature_adapt_cls.conv_adaption.weight", "tasks.0.feature_adapt_reg.conv_offset.weight",
         "tasks.0.feature_adapt_reg.conv_offset.bias", "tasks.0.feature_adapt_reg.conv_adaption.weight",
         "tasks.0.cls_head.0.weight", "tasks.0.cls_head.0.bias", "tasks.0.cls_head.1.weight", "tasks.0.cls_head.1.bias",
         "tasks.0.cls_head.3.weight", "tasks.0.cls_head.3.bias", "tasks.0.task_head.reg.0.weight", "tasks.0.task_head.reg.0.bias",
         "tasks.0.task_head.reg.2.weight", "tasks.0.task_head.reg.2.bias"])
