"""torch.autograd Functions over NCDHW fp32 tensors that run on the P8 tensor-core kernels: conv3d (k = 1 or 3, stride 1,
'same' padding, optional bias / fused ReLU) and GroupNorm.  They are the building blocks of module-level heads that are
not part of the fused engine program — the 3-D-compatible deformable head (det3d_compat.DCNSepHead) — so that those
heads, too, execute only librtpose_b200.so kernels for their contractions and normalisations (the fp32 <-> P8 boundary
conversions are rtp_pack_ncdhw / rtp_unpack_ncdhw).

Reference ops they stand in for: nn.Conv3d / nn.Conv2d (center_head.py:86-93, :134-142), nn.GroupNorm (:85, :204).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import lib, ops
from .p8 import P8

_packs = ops.PackedWeights()


def _cuda(x, what):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise lib.RtpError("%s must be a CUDA tensor: the rtpose_b200 path has no CPU fallback" % what)
    return x


class _Conv3d(Function):
    @staticmethod
    def forward(ctx, x, w, bias, relu):
        x = _cuda(x, "conv3d input").contiguous().float()
        N, Cin, Z, Y, X = x.shape
        Cout, k = w.shape[0], w.shape[2]
        if w.shape[1] != Cin or k not in (1, 3) or tuple(w.shape[2:]) != (k, k, k):
            raise lib.RtpError("conv3d: weight %s does not fit input %s (k must be 1 or 3)" % (tuple(w.shape), tuple(x.shape)))
        xp = P8.from_ncdhw(x)
        yp = P8(N, Cout, Z, Y, X, device=x.device)
        ops.conv_forward(_packs, xp, w, 1, yp, bias=bias, relu=bool(relu))
        ctx.xp, ctx.yp, ctx.w, ctx.relu, ctx.has_bias = xp, (yp if relu else None), w, bool(relu), bias is not None
        return yp.to_ncdhw()

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xp, w, k = ctx.xp, ctx.w, ctx.w.shape[2]
        dy = P8.from_ncdhw(gy.contiguous().float())
        if ctx.relu:  # dL/d(pre-ReLU) = dy * (y > 0)
            masked = P8(dy.N, dy.C, dy.Z, dy.Y, dy.X, device=gy.device)
            ops.grad_add(dy, masked, mask=ctx.yp)
            dy = masked
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            dx = P8(xp.N, xp.C, xp.Z, xp.Y, xp.X, device=gy.device)
            ops.conv_dgrad(_packs, dy, w, 1, dx)
            gx = dx.to_ncdhw()
        if ctx.needs_input_grad[1]:
            gw = torch.empty(w.shape, dtype=torch.float32, device=gy.device)
            ops.conv_wgrad(xp, dy, k, 1, gw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = torch.empty(w.shape[0], dtype=torch.float32, device=gy.device)
            ops.channel_sum(dy, gb)
        ctx.xp = ctx.yp = None
        return gx, gw, gb, None


def conv3d(x, weight, bias=None, relu=False):
    """y = [relu](conv3d(x, weight, bias, stride=1, padding=k//2)); x fp32 [N,C,Z,Y,X] on CUDA, weight [Cout,Cin,k,k,k]."""
    return _Conv3d.apply(x, weight, bias, relu)


def conv2d_as_3d(x5, weight2d, bias=None, relu=False):
    """A 2-D k x k conv applied to every z-slice of x5 [N,C,Z,Y,X] (the reference's Conv2d on the z-folded batch,
    center_head.py:134-142): the 2-D kernel sits in the middle z-plane of a k x k x k kernel whose other planes are zero, so
    the plane-streaming kernel computes it; autograd slices the weight gradient back."""
    Cout, Cin, kh, kw = weight2d.shape
    if kh != kw or kh not in (1, 3):
        raise lib.RtpError("conv2d_as_3d: kernel %dx%d" % (kh, kw))
    if kh == 1:
        return conv3d(x5, weight2d.reshape(Cout, Cin, 1, 1, 1), bias, relu)
    w3 = torch.zeros((Cout, Cin, 3, 3, 3), dtype=weight2d.dtype, device=weight2d.device)
    w3[:, :, 1] = weight2d
    return conv3d(x5, w3, bias, relu)


class _GroupNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, G):
        x = _cuda(x, "group_norm input").contiguous().float()
        xp = P8.from_ncdhw(x)
        stats = ops.gn_stats(xp, G)
        yp = ops.gn_apply(xp, G, stats, gamma, beta, P8(xp.N, xp.C, xp.Z, xp.Y, xp.X, device=x.device))
        ctx.xp, ctx.stats, ctx.gamma, ctx.G = xp, stats, gamma, G
        return yp.to_ncdhw()

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xp, G = ctx.xp, ctx.G
        dy = P8.from_ncdhw(gy.contiguous().float())
        dg = torch.empty(xp.C, dtype=torch.float32, device=gy.device)
        db = torch.empty(xp.C, dtype=torch.float32, device=gy.device)
        dx = P8(xp.N, xp.C, xp.Z, xp.Y, xp.X, device=gy.device)
        ops.gn_backward(xp, dy, G, ctx.stats, ctx.gamma, dg, db, False, dx, False)
        ctx.xp = None
        return dx.to_ncdhw(), dg, db, None


def group_norm(x, num_groups, gamma, beta):
    """GroupNorm(num_groups, C) (eps 1e-5, biased variance) of x fp32 [N,C,Z,Y,X]."""
    return _GroupNorm.apply(x, gamma, beta, num_groups)
