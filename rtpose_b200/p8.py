"""P8 activation tensors: bf16, 8-channel blocked, in-plane zero-padded, y fastest — see include/rtpose_b200.h.

A `P8` owns (or views) a flat torch bf16 buffer with guard bytes on both sides; pads are zeroed once at
allocation and no kernel ever writes them.  torch is used for device memory only.
"""
import torch

from . import lib

GUARD_ELEMS = lib.GUARD_BYTES // 2


class P8:
    __slots__ = ("buf", "offset", "N", "C", "C8", "Z", "Y", "X", "n_stride", "c_stride", "relu_out", "grad", "grad_ev", "pending", "_keep")

    def __init__(self, N, C, Z, Y, X, device="cuda", buf=None, offset=None, n_stride=None, c_stride=None):
        self.N, self.C, self.Z, self.Y, self.X = int(N), int(C), int(Z), int(Y), int(X)
        self.C8 = (self.C + 7) // 8
        plane = (self.X + 2) * (self.Y + 2) * 8
        self.c_stride = int(c_stride) if c_stride is not None else self.Z * plane
        self.n_stride = int(n_stride) if n_stride is not None else self.C8 * self.c_stride
        if buf is None:
            total = self.N * self.n_stride + 2 * GUARD_ELEMS
            buf = torch.zeros(total, dtype=torch.bfloat16, device=device)
            offset = GUARD_ELEMS
        self.buf = buf
        self.offset = int(offset)
        self.relu_out = False  # True when the tensor is the output of a ReLU (gradients are kept pre-ReLU)
        self.grad = None
        self.grad_ev = None  # event of the last kernel that wrote .grad (ordered accumulation across streams)
        self.pending = None  # a gradient still to be added into .grad (engine._defer_add): folded into the next GroupNorm backward
        self._keep = None

    # ------------------------------------------------------------------ views
    @property
    def ptr(self):
        return self.buf.data_ptr() + 2 * self.offset

    def struct(self):
        return lib.P8Struct(self.ptr, self.n_stride, self.c_stride, self.N, self.C8, self.Z, self.X, self.Y)

    def channels(self, c0, c):
        """View of channels [c0, c0+c) (c0 multiple of 8)."""
        assert c0 % 8 == 0
        v = P8(self.N, c, self.Z, self.Y, self.X, buf=self.buf, offset=self.offset + (c0 // 8) * self.c_stride,
               n_stride=self.n_stride, c_stride=self.c_stride)
        v.relu_out = self.relu_out
        return v

    @property
    def device(self):
        return self.buf.device

    def like(self, C=None):
        return P8(self.N, self.C if C is None else C, self.Z, self.Y, self.X, device=self.buf.device)

    @property
    def voxels(self):
        return self.Z * self.Y * self.X

    @property
    def grid(self):
        return (self.Z, self.Y, self.X)

    # ------------------------------------------------------------------ boundary conversion
    @staticmethod
    def from_ncdhw(x):
        """fp32 [N,C,Z,Y,X] CUDA tensor -> P8 (kernel rtp_pack_ncdhw)."""
        assert x.is_cuda and x.dim() == 5
        x = x.contiguous().float()
        N, Cc, Z, Y, X = x.shape
        t = P8(N, Cc, Z, Y, X, device=x.device)
        lib.call("rtp_pack_ncdhw", x.data_ptr(), t.struct(), Cc, _stream())
        return t

    def to_ncdhw(self, out=None, accumulate=False):
        """P8 -> fp32 [N,C,Z,Y,X] (kernel rtp_unpack_ncdhw)."""
        if out is None:
            out = torch.empty((self.N, self.C, self.Z, self.Y, self.X), dtype=torch.float32, device=self.buf.device)
            accumulate = False
        lib.call("rtp_unpack_ncdhw", self.struct(), out.data_ptr(), self.C, 1 if accumulate else 0, _stream())
        return out

    def zero_(self):
        if self.n_stride == self.C8 * self.c_stride:
            self.buf[self.offset:self.offset + self.N * self.n_stride].zero_()
        else:
            self.buf.as_strided((self.N, self.C8 * self.c_stride), (self.n_stride, 1), self.offset).zero_()
        return self


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream():
    """Raw cudaStream_t of torch's current stream (torch.cuda.current_stream() costs ~10 us of Python per call)."""
    return _raw_stream(_cur_device())
