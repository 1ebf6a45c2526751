"""Deformable convolution v1 and v2 (2-D) with the reference's Python API (det3d/ops/dcn/__init__.py: DeformConv,
DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack, deform_conv, modulated_deform_conv).

Mirrors det3d/ops/dcn/deform_conv.py:14-112 (`DeformConvFunction`: forward / backward, 8-argument signature,
`_output_size`) and :192-255 (`DeformConv` module: no bias, uniform(-1/sqrt(fan_in)) init, small-input padding
work-around).  The kernels are librtpose_b200.so's rtp_dcn_* (no `columns` buffer, no im2col_step batching — the
argument is accepted and ignored).  groups must be 1, as in every use the reference makes of the op
(center_head.py:45-51: DeformConv(C, C, 3, padding=1, deformable_groups=4)).

v2: `ModulatedDeformConvFunction` (deform_conv.py:115-186), `ModulatedDeformConv` (:326-379) and the two `*Pack` modules
(:258-323, :382-446), whose offset/mask predictor is an ordinary nn.Conv2d exactly as in the reference; kernels rtp_mdcn_*.
"""
import math
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair, _single

from . import lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Forward contraction on the tcgen05 tensor cores (bf16 operands, fp32 accumulation) instead of the fp32 CUDA-core kernel:
# the deformed samples are written once as a bf16 P8 volume with the taps on the z axis (rtp_dcn_sample_p8) and contracted
# by rtp_conv.  This is the DEFAULT for every shape it supports (tc_supported): the op then computes in bf16-in / fp32-
# accumulate like the rest of the path, 2-4x faster than the fp32 kernels (profiles/r01_dcn_timings.txt).  RTP_DCN_TC=0 or
# `rtpose_b200.dcn.TENSOR_CORE = False` selects the fp32 CUDA-core kernels (bit-for-bit the reference op's precision).
TENSOR_CORE = os.environ.get("RTP_DCN_TC", "1") not in ("", "0")
# The same for the backward pass (RTP_DCN_TC_BWD=1 / TENSOR_CORE_BACKWARD): weight gradient = rtp_wgrad over (sample
# volume, dy), sample gradient = kh*kw single-tap rtp_conv launches with the dgrad-packed weight, then rtp_dcn_col2im_p8
# scatters it into dx / doffset / dmask.  Gradients then carry bf16 operand rounding like every other conv of the path.
TENSOR_CORE_BACKWARD = os.environ.get("RTP_DCN_TC_BWD", "1") not in ("", "0")
TC_SAMPLE_BYTES = 2 << 30     # the sampled volume is produced and consumed in batch chunks of at most this size
_tc_state = {}


def _same_tensors(entry, tensors):
    """A cached pack is valid only for the SAME tensor objects at the same version counters: a new tensor that happens to
    reuse a freed tensor's address (and starts at version 0 as well) must not hit."""
    if entry is None:
        return False
    refs, vers = entry
    return all((r is None and t is None) or (r is not None and t is not None and r() is t and v == t._version)
               for r, v, t in zip(refs, vers, tensors))


def _tensor_key(tensors):
    import weakref
    return ([weakref.ref(t) if t is not None else None for t in tensors], [t._version if t is not None else None for t in tensors])


def tc_supported(C, Cout, kh, kw, dg):
    return (C % dg == 0 and (C // dg) % 8 == 0 and -(-C // 16) * 16 <= 512 and -(-Cout // 16) * 16 <= 256 and kh * kw <= lib.MAX_TAPS)


def _tc_chunk(N, Cc, K, Ho, Wo, dg):
    per = (-(-Cc // 8)) * K * (Wo + 2) * (Ho + 2) * 16
    return max(1, min(N, TC_SAMPLE_BYTES // per, 65535 // (dg * K)))


def _tc_backward(input, offset, mask, weight, grad_output, with_bias, stride, pad, dil, dg):
    """(grad_input, grad_offset, grad_mask | None, grad_weight, grad_bias | None) on the tensor-core path."""
    from . import ops
    from .p8 import P8
    N, Cc, H, W = input.shape
    Cout, _, kh, kw = weight.shape
    K, Ho, Wo = kh * kw, grad_output.shape[2], grad_output.shape[3]
    KPd, NPd = -(-Cout // 16) * 16, -(-Cc // 16) * 16  # dgrad GEMM: K = Cout, N = C
    dev = input.device
    nb = _tc_chunk(N, Cc, K, Ho, Wo, dg)
    key = (dev, nb, Cc, K, Ho, Wo, Cout, "bwd")
    st = _tc_state.get(key)
    if st is None:
        st = _tc_state[key] = {"S": P8(nb, Cc, K, Ho, Wo, device=dev), "w": None}
    if st["w"] is None or not _same_tensors(st["w"][0], (weight,)):
        pack = torch.empty(K * KPd * NPd, dtype=torch.bfloat16, device=dev)
        lib.call("rtp_weight_pack", weight.data_ptr(), pack.data_ptr(), Cout, Cc, K, 0, Cc, KPd, NPd, 1, _stream())
        st["w"] = (_tensor_key((weight,)), pack)
    pack = st["w"][1]
    grad_input, grad_offset = torch.empty_like(input), torch.empty_like(offset)
    grad_mask = torch.empty_like(mask) if mask is not None else None
    grad_weight = torch.empty_like(weight)
    taps = [(t, 0, 0) for t in range(K)]
    for n0 in range(0, N, nb):
        n = min(nb, N - n0)
        S = st["S"] if n == nb else P8(n, Cc, K, Ho, Wo, buf=st["S"].buf, offset=st["S"].offset)
        mptr = mask[n0:].data_ptr() if mask is not None else None
        lib.call("rtp_dcn_sample_p8", input[n0:].data_ptr(), offset[n0:].data_ptr(), mptr, S.struct(), n, Cc, H, W, kh, kw, stride, pad,
                 dil, dg, _stream())
        dY = P8.from_ncdhw(grad_output[n0:n0 + n].reshape(n, Cout, 1, Ho, Wo))
        ops.conv_wgrad(S, dY, 1, 1, grad_weight, accumulate=n0 > 0, taps=taps)
        for t in range(K):  # the sample volume is dead now: its buffer receives the sample gradient, one tap plane per launch
            ops.conv(dY, pack, KPd, NPd, S, [(0, 0, 0, t)], (1, Wo, Ho), off=(t, 0, 0), real=(Cout, Cc))
        lib.call("rtp_dcn_col2im_p8", input[n0:].data_ptr(), offset[n0:].data_ptr(), mptr, S.struct(), grad_input[n0:].data_ptr(),
                 grad_offset[n0:].data_ptr(), grad_mask[n0:].data_ptr() if mask is not None else None, n, Cc, H, W, kh, kw, stride, pad,
                 dil, dg, _stream())
    grad_bias = None
    if with_bias:
        grad_bias = torch.zeros(Cout, dtype=torch.float32, device=dev)
        lib.call("rtp_dcn_bias_grad", grad_output.data_ptr(), grad_bias.data_ptr(), N, Cout, Ho * Wo, 1.0, _stream())
    return grad_input, grad_offset, grad_mask, grad_weight, grad_bias


def _tc_forward(input, offset, mask, weight, bias, stride, pad, dil, dg, out_size):
    """input fp32 [N,C,H,W] -> output fp32 [N,Cout,Ho,Wo] through rtp_dcn_sample_p8 + rtp_conv (+ rtp_unpack_ncdhw)."""
    from . import ops
    from .p8 import P8
    N, Cc, H, W = input.shape
    Cout, _, kh, kw = weight.shape
    K, Ho, Wo = kh * kw, out_size[2], out_size[3]
    if not tc_supported(Cc, Cout, kh, kw, dg):
        raise lib.RtpError("tensor-core DCN needs (C/dg) %% 8 == 0, C <= 512, Cout <= 256 (got C=%d Cout=%d dg=%d)" % (Cc, Cout, dg))
    KP, NP = -(-Cc // 16) * 16, -(-Cout // 16) * 16
    nb = _tc_chunk(N, Cc, K, Ho, Wo, dg)
    dev = input.device
    key = (dev, nb, Cc, K, Ho, Wo, Cout)
    st = _tc_state.get(key)
    if st is None:  # zero-filled once: the kernels only ever write the interior, the halo stays zero
        st = _tc_state[key] = {"S": P8(nb, Cc, K, Ho, Wo, device=dev), "Y": P8(nb, Cout, 1, Ho, Wo, device=dev), "w": None}
    if st["w"] is None or not _same_tensors(st["w"][0], (weight, bias)):
        pack = torch.empty(K * KP * NP, dtype=torch.bfloat16, device=dev)
        lib.call("rtp_weight_pack", weight.data_ptr(), pack.data_ptr(), Cout, Cc, K, 0, Cc, KP, NP, 0, _stream())
        st["w"] = (_tensor_key((weight, bias)), pack, ops.pad_bias(bias, NP) if bias is not None else None)
    _, pack, bias_p = st["w"]
    taps = [(t, 0, 0, t) for t in range(K)]
    out = input.new_empty(out_size)
    for n0 in range(0, N, nb):
        n = min(nb, N - n0)
        S, Y = st["S"], st["Y"]
        if n != nb:
            S, Y = P8(n, Cc, K, Ho, Wo, buf=S.buf, offset=S.offset), P8(n, Cout, 1, Ho, Wo, buf=Y.buf, offset=Y.offset)
        lib.call("rtp_dcn_sample_p8", input[n0:].data_ptr(), offset[n0:].data_ptr(), mask[n0:].data_ptr() if mask is not None else None,
                 S.struct(), n, Cc, H, W, kh, kw, stride, pad, dil, dg, _stream())
        ops.conv(S, pack, KP, NP, Y, taps, (1, Wo, Ho), bias=bias_p, real=(Cc, Cout))
        lib.call("rtp_unpack_ncdhw", Y.struct(), out[n0:].data_ptr(), Cout, 0, _stream())
    return out


class DeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        if not input.is_cuda:
            raise NotImplementedError  # same as the reference (deform_conv.py:46-47)
        if groups != 1:
            raise NotImplementedError("rtpose_b200 DeformConv supports groups=1 only")
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        if ctx.stride[0] != ctx.stride[1] or ctx.padding[0] != ctx.padding[1] or ctx.dilation[0] != ctx.dilation[1]:
            raise NotImplementedError("anisotropic stride/padding/dilation")
        ctx.groups, ctx.deformable_groups, ctx.im2col_step = groups, deformable_groups, im2col_step
        input, offset, weight = input.contiguous().float(), offset.contiguous().float(), weight.contiguous().float()
        ctx.save_for_backward(input, offset, weight)
        out_size = DeformConvFunction._output_size(input, weight, ctx.padding, ctx.dilation, ctx.stride)
        N, Cc, H, W = input.shape
        if offset.shape != (N, deformable_groups * 2 * weight.shape[2] * weight.shape[3], out_size[2], out_size[3]):
            raise ValueError("invalid offset shape {} for output {}".format(tuple(offset.shape), out_size))
        if TENSOR_CORE and tc_supported(Cc, weight.shape[0], weight.shape[2], weight.shape[3], deformable_groups):
            return _tc_forward(input, offset, None, weight, None, ctx.stride[0], ctx.padding[0], ctx.dilation[0], deformable_groups,
                               out_size)
        output = input.new_empty(out_size)
        lib.call("rtp_dcn_fwd", input.data_ptr(), offset.data_ptr(), weight.data_ptr(), output.data_ptr(), N, Cc, H, W,
                 weight.shape[0], weight.shape[2], weight.shape[3], ctx.stride[0], ctx.padding[0], ctx.dilation[0],
                 deformable_groups, _stream())
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        N, Cc, H, W = input.shape
        args = (N, Cc, H, W, weight.shape[0], weight.shape[2], weight.shape[3], ctx.stride[0], ctx.padding[0],
                ctx.dilation[0], ctx.deformable_groups)
        grad_input = grad_offset = grad_weight = None
        if TENSOR_CORE_BACKWARD and tc_supported(Cc, weight.shape[0], weight.shape[2], weight.shape[3], ctx.deformable_groups):
            grad_input, grad_offset, _, grad_weight, _ = _tc_backward(input, offset, None, weight, grad_output, False, ctx.stride[0],
                                                                      ctx.padding[0], ctx.dilation[0], ctx.deformable_groups)
            return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_input, grad_offset = torch.empty_like(input), torch.empty_like(offset)
            lib.call("rtp_dcn_bwd_input", input.data_ptr(), offset.data_ptr(), weight.data_ptr(), grad_output.data_ptr(),
                     grad_input.data_ptr(), grad_offset.data_ptr(), *args, _stream())
        if ctx.needs_input_grad[2]:
            grad_weight = torch.zeros_like(weight)
            lib.call("rtp_dcn_bwd_weight", input.data_ptr(), offset.data_ptr(), grad_output.data_ptr(), grad_weight.data_ptr(),
                     *args, 1.0, _stream())
        return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)

    @staticmethod
    def _output_size(input, weight, padding, dilation, stride):
        channels = weight.size(0)
        output_size = (input.size(0), channels)
        for d in range(input.dim() - 2):
            in_size = input.size(d + 2)
            pad = padding[d]
            kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
            output_size += ((in_size + (2 * pad) - kernel) // stride[d] + 1,)
        if not all(map(lambda s: s > 0, output_size)):
            raise ValueError("convolution input is too small (output would be {})".format("x".join(map(str, output_size))))
        return output_size


deform_conv = DeformConvFunction.apply


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super(DeformConv, self).__init__()
        assert not bias
        assert in_channels % groups == 0, "in_channels {} cannot be divisible by groups {}".format(in_channels, groups)
        assert out_channels % groups == 0, "out_channels {} cannot be divisible by groups {}".format(out_channels, groups)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.transposed, self.output_padding = False, _single(0)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        input_pad = x.size(2) < self.kernel_size[0] or x.size(3) < self.kernel_size[1]
        if input_pad:
            pad_h = max(self.kernel_size[0] - x.size(2), 0)
            pad_w = max(self.kernel_size[1] - x.size(3), 0)
            x = F.pad(x, (0, pad_w, 0, pad_h), "constant", 0).contiguous()
            offset = F.pad(offset, (0, pad_w, 0, pad_h), "constant", 0).contiguous()
        out = deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups, self.deformable_groups)
        if input_pad:
            out = out[:, :, :out.size(2) - pad_h, :out.size(3) - pad_w].contiguous()
        return out


class DeformConvPack(DeformConv):
    """det3d/ops/dcn/deform_conv.py:258-323: DeformConv that predicts its own offsets with a zero-initialised conv."""

    _version = 2

    def __init__(self, *args, **kwargs):
        super(DeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        return deform_conv(x, self.conv_offset(x), self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        _rename_v1_offset_keys(state_dict, prefix, local_metadata)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def _rename_v1_offset_keys(state_dict, prefix, local_metadata):
    """Checkpoints written before module version 2 call the predictor `<name>_offset` instead of `<name>.conv_offset`."""
    version = local_metadata.get("version", None)
    if version is None or version < 2:
        for leaf in ("weight", "bias"):
            new, old = prefix + "conv_offset." + leaf, prefix[:-1] + "_offset." + leaf
            if new not in state_dict and old in state_dict:
                state_dict[new] = state_dict.pop(old)


class ModulatedDeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
        if not input.is_cuda:
            raise NotImplementedError  # deform_conv.py:137-138
        if input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        if groups != 1:
            raise NotImplementedError("rtpose_b200 ModulatedDeformConv supports groups=1 only")
        if any(isinstance(v, (tuple, list)) and v[0] != v[1] for v in (stride, padding, dilation)):
            raise NotImplementedError("anisotropic stride/padding/dilation")
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride)[0], _pair(padding)[0], _pair(dilation)[0]
        ctx.groups, ctx.deformable_groups, ctx.with_bias = groups, deformable_groups, bias is not None
        input, offset, mask, weight = (t.contiguous().float() for t in (input, offset, mask, weight))
        bias = bias.contiguous().float() if ctx.with_bias else None
        N, Cc, H, W = input.shape
        out_size = ModulatedDeformConvFunction._infer_shape(ctx, input, weight)
        K = weight.shape[2] * weight.shape[3]
        if offset.shape != (N, deformable_groups * 2 * K, out_size[2], out_size[3]):
            raise ValueError("invalid offset shape {} for output {}".format(tuple(offset.shape), out_size))
        if mask.shape != (N, deformable_groups * K, out_size[2], out_size[3]):
            raise ValueError("invalid mask shape {} for output {}".format(tuple(mask.shape), out_size))
        ctx.save_for_backward(input, offset, mask, weight)
        if TENSOR_CORE and tc_supported(Cc, weight.shape[0], weight.shape[2], weight.shape[3], deformable_groups):
            return _tc_forward(input, offset, mask, weight, bias, ctx.stride, ctx.padding, ctx.dilation, deformable_groups, out_size)
        output = input.new_empty(out_size)
        lib.call("rtp_mdcn_fwd", input.data_ptr(), offset.data_ptr(), mask.data_ptr(), weight.data_ptr(),
                 bias.data_ptr() if ctx.with_bias else None, output.data_ptr(), N, Cc, H, W, weight.shape[0], weight.shape[2],
                 weight.shape[3], ctx.stride, ctx.padding, ctx.dilation, deformable_groups, _stream())
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        N, Cc, H, W = input.shape
        args = (N, Cc, H, W, weight.shape[0], weight.shape[2], weight.shape[3], ctx.stride, ctx.padding, ctx.dilation,
                ctx.deformable_groups)
        if TENSOR_CORE_BACKWARD and tc_supported(Cc, weight.shape[0], weight.shape[2], weight.shape[3], ctx.deformable_groups):
            return _tc_backward(input, offset, mask, weight, grad_output, ctx.with_bias, ctx.stride, ctx.padding, ctx.dilation,
                                ctx.deformable_groups) + (None, None, None, None, None)
        grad_input, grad_offset, grad_mask = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
        lib.call("rtp_mdcn_bwd_input", input.data_ptr(), offset.data_ptr(), mask.data_ptr(), weight.data_ptr(),
                 grad_output.data_ptr(), grad_input.data_ptr(), grad_offset.data_ptr(), grad_mask.data_ptr(), *args, _stream())
        grad_weight = torch.zeros_like(weight)
        grad_bias = torch.zeros(weight.shape[0], dtype=torch.float32, device=weight.device) if ctx.with_bias else None
        lib.call("rtp_mdcn_bwd_weight", input.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_output.data_ptr(),
                 grad_weight.data_ptr(), grad_bias.data_ptr() if ctx.with_bias else None, *args, 1.0, _stream())
        return (grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None)

    @staticmethod
    def _infer_shape(ctx, input, weight):
        n, channels_out = input.size(0), weight.size(0)
        height, width = input.shape[2:4]
        kernel_h, kernel_w = weight.shape[2:4]
        height_out = (height + 2 * ctx.padding - (ctx.dilation * (kernel_h - 1) + 1)) // ctx.stride + 1
        width_out = (width + 2 * ctx.padding - (ctx.dilation * (kernel_w - 1) + 1)) // ctx.stride + 1
        return n, channels_out, height_out, width_out


modulated_deform_conv = ModulatedDeformConvFunction.apply


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                 bias=True):
        super(ModulatedDeformConv, self).__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups, self.with_bias = groups, deformable_groups, bias
        self.transposed, self.output_padding = False, _single(0)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                                     self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """det3d/ops/dcn/deform_conv.py:382-446: predicts 2K offset + K mask channels per deformable group with one
    zero-initialised conv; mask = sigmoid of the last third, offset = the first two thirds concatenated."""

    _version = 2

    def __init__(self, *args, **kwargs):
        super(ModulatedDeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        o1, o2, mask = torch.chunk(self.conv_offset(x), 3, dim=1)
        return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        _rename_v1_offset_keys(state_dict, prefix, local_metadata)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class FeatureAdaption(nn.Module):
    """det3d/models/pose_heads/center_head.py:24-62: offsets from a zero-weight 1x1 conv (its bias keeps the default
    init, as in the reference), DCN v1 with `deformable_groups` groups, ReLU.  state_dict keys: conv_offset.{weight,bias},
    conv_adaption.weight."""

    def __init__(self, in_channels, out_channels, kernel_size=3, deformable_groups=4):
        super(FeatureAdaption, self).__init__()
        offset_channels = kernel_size * kernel_size * 2
        self.conv_offset = nn.Conv2d(in_channels, deformable_groups * offset_channels, 1, bias=True)
        self.conv_adaption = DeformConv(in_channels, out_channels, kernel_size=kernel_size, padding=(kernel_size - 1) // 2,
                                        deformable_groups=deformable_groups)
        self.relu = nn.ReLU(inplace=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()

    def forward(self, x):
        offset = self.conv_offset(x)
        return self.relu(self.conv_adaption(x, offset))
