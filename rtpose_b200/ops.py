"""Host-side wrappers of the C-ABI kernels: descriptor construction (tap lists, row grids), weight-pack caching,
workspace and activation-buffer pooling.  Everything here launches on torch's current CUDA stream.
"""
import ctypes as C
import os as _os

import torch

from . import lib
from .p8 import P8, _stream


def ceil_to(v, m):
    return (v + m - 1) // m * m


# ------------------------------------------------------------------------------------------------ pools
class BufferPool:
    """Recycles P8 buffers between steps.  Pads are zeroed at first allocation and never written afterwards, and
    every producer kernel overwrites all real voxels, so recycled buffers need no re-zeroing."""

    def __init__(self):
        self.free = {}
        self.live = []

    def get(self, N, Cc, Z, Y, X, device):
        key = (N, (Cc + 7) // 8, Z, Y, X, str(device))
        lst = self.free.get(key)
        if lst:
            t = lst.pop()
            t.C = Cc
            t.relu_out = False
            t.grad = None
            t.grad_ev = None
            t.pending = None
        else:
            t = P8(N, Cc, Z, Y, X, device=device)
        t._keep = key
        self.live.append(t)
        return t

    def release_all(self):
        for t in self.live:
            t.grad = None
            t.grad_ev = None
            t.pending = None
            self.free.setdefault(t._keep, []).append(t)
        self.live = []


_workspaces = {}
_streams = {}


def named_stream(device, name, priority=0):
    """One CUDA stream per (device, role) for the whole process.  torch hands out streams from a pool of 32 per priority and
    device, round-robin: code that creates fresh streams per engine / per graph capture starts to ALIAS them once the pool
    wraps (two roles on one cudaStream_t), which showed up as an order-dependent capture invalidation in the test-suite.
    Engines, the weight-gradient / pack / communication streams and StepGraph therefore share these few named streams."""
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (str(dev), name)
    st = _streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev, priority=priority)
        _streams[key] = st
    return st



def workspace(nbytes, device, tag="ws"):
    """Scratch buffer per (tag, device, current stream): kernels of concurrently running streams never share one."""
    key = (tag, str(device), _stream())
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


# ------------------------------------------------------------------------------------------------ tap lists
def taps_fwd(k):
    """(tz, tx, ty, wt) for a forward conv; wt = (kz*k + ky)*k + kx indexes the reference weight's taps."""
    if k == 1:
        return [(0, 0, 0, 0)]
    return [(kz - 1, kx - 1, ky - 1, (kz * 3 + ky) * 3 + kx) for kz in range(3) for ky in range(3) for kx in range(3)]


def taps_dgrad_s1(k):
    if k == 1:
        return [(0, 0, 0, 0)]
    return [(1 - kz, 1 - kx, 1 - ky, (kz * 3 + ky) * 3 + kx) for kz in range(3) for ky in range(3) for kx in range(3)]


def taps_dgrad_s2(pz, px, py):
    """dgrad of a 3x3x3 stride-2 pad-1 conv for input voxels of parity (pz, px, py): input index i = 2*o + k - 1, so
    even i pairs with k=1 (o = i/2) and odd i with k=0 (o = (i+1)/2) or k=2 (o = (i-1)/2)."""
    def dim(p):
        return [(1, 0)] if p == 0 else [(0, 1), (2, 0)]
    return [(tz, tx, ty, (kz * 3 + ky) * 3 + kx) for kz, tz in dim(pz) for ky, ty in dim(py) for kx, tx in dim(px)]


def _fill_taps(d, taps, with_wt=True):
    for i, t in enumerate(taps):
        d.tz[i], d.tx[i], d.ty[i] = t[0], t[1], t[2]
        if with_wt:
            d.wt[i] = t[3]
    d.ntaps = len(taps)


# ------------------------------------------------------------------------------------------------ weights
class PackedWeights:
    """bf16 UMMA-B packs of an fp32 conv weight, rebuilt when the parameter's version counter changes.

    A cache entry is valid only for the SAME tensor object (weak reference) at the same version — a new tensor that
    happens to reuse a freed tensor's address never hits.  Temporaries (merged / sliced weights) must pass an
    explicit `key` and `version` derived from the parameters they were built from."""

    def __init__(self):
        self.cache = {}      # key -> (version, packed value, weakref to the source weight, repack(w) or None)
        self.jobs = {}       # key -> job of the batched repack (entries with a live, contiguous source weight)
        self._batch = None   # (signature, device job table, njobs, total blocks) of the last refresh_async()
        self._side = None    # stream of a refresh_async() still to be joined by the first consumer

    def _lookup(self, key, w, ver, explicit):
        if self._side is not None:  # the first consumer after refresh_async() joins the packing stream
            torch.cuda.current_stream(w.device).wait_stream(self._side)
            self._side = None
        hit = self.cache.get(key)
        if hit is None:
            return None, None
        hver, val, ref, _ = hit
        same = explicit or (ref is not None and ref() is w)
        return (val if (same and hver == ver and hver is not None) else None), val

    def _store(self, key, w, ver, val, explicit, repack, job=None):
        import weakref
        ref = None
        if not explicit:
            try:
                ref = weakref.ref(w)
            except TypeError:
                ref = None
        self.cache[key] = (ver, val, ref, repack if ref is not None else None)
        if ref is not None and job is not None and w.is_contiguous():
            self.jobs[key] = job  # (kind, dst tensor, Cout, Cin_total, ntaps, ci0, ci_n, KP, NP, flag): rtp_weight_pack_batch
        else:
            self.jobs.pop(key, None)

    def get(self, w, mode, ci0=0, ci_n=None, key=None, version=None):
        Cout, Cin = w.shape[0], w.shape[1]
        ntaps = w.shape[2] * w.shape[3] * w.shape[4]
        ci_n = Cin if ci_n is None else ci_n
        explicit = key is not None
        key = (key if explicit else w.data_ptr(), mode, ci0, ci_n, tuple(w.shape))
        ver = version if version is not None else w._version
        val, old = self._lookup(key, w, ver, explicit)
        if val is not None:
            return val
        if mode == 0:
            KP, NP = ceil_to(ci_n, 16), ceil_to(Cout, 16)
        else:
            KP, NP = ceil_to(Cout, 16), ceil_to(ci_n, 16)
        dst = old[0] if old is not None else torch.empty(ntaps * KP * NP, dtype=torch.bfloat16, device=w.device)

        def repack(src):
            wc = src.detach().contiguous()
            lib.call("rtp_weight_pack", wc.data_ptr(), dst.data_ptr(), Cout, Cin, ntaps, ci0, ci_n, KP, NP, mode, _stream())
        repack(w)
        val = (dst, KP, NP)
        self._store(key, w, ver, val, explicit, repack, (0, dst, Cout, Cin, ntaps, ci0, ci_n, KP, NP, mode))
        return val

    def get_k3s1(self, w, K, NPo, transpose_flip, key=None, version=None, ci_window=None):
        """kz-stacked pack for the plane-streaming kernel: [9][K/8][3*NPo][8].  ci_window = (ci0, ci_n): pack only that
        window of w's input channels (w stays the full, contiguous weight: no slice is materialised)."""
        explicit = key is not None
        key = (key if explicit else w.data_ptr(), "k3s1", bool(transpose_flip), K, NPo, tuple(w.shape), ci_window)
        ver = version if version is not None else w._version
        val, old = self._lookup(key, w, ver, explicit)
        if val is not None:
            return val
        dst = old if old is not None else torch.empty(9 * K * 3 * NPo, dtype=torch.bfloat16, device=w.device)
        shape = (w.shape[0], w.shape[1])

        def repack(src):
            wc = src.detach().contiguous()
            if ci_window is not None:
                lib.call("rtp_weight_pack_k3s1_window", wc.data_ptr(), dst.data_ptr(), shape[0], shape[1], ci_window[0], ci_window[1],
                         K, NPo, int(bool(transpose_flip)), _stream())
                return
            lib.call("rtp_weight_pack_k3s1", wc.data_ptr(), dst.data_ptr(), shape[0], shape[1], K, NPo,
                     int(bool(transpose_flip)), _stream())
        repack(w)
        cw = ci_window if ci_window is not None else (0, shape[1])
        self._store(key, w, ver, dst, explicit, repack, (1, dst, shape[0], shape[1], 27, cw[0], cw[1], K, NPo, int(bool(transpose_flip))))
        return dst

    def refresh_async(self):
        """Training: the optimizer has rewritten the weights in place.  Repacks every cached weight that still has a live
        source tensor on a side stream (forked after the work queued so far; the first consumer joins it), instead of
        lazily, one small kernel at a time, in front of each conv.  Entries built from per-step temporaries (explicit
        keys) are only invalidated."""
        dev = None
        live = []
        for k in list(self.cache):
            ver, val, ref, repack = self.cache[k]
            w = ref() if ref is not None else None
            if repack is None or w is None:
                if ref is not None and w is None:
                    del self.cache[k]          # the source tensor is gone
                    self.jobs.pop(k, None)
                else:
                    self.cache[k] = (None, val, ref, repack)
                continue
            if dev is None:
                dev = w.device
                global _pack_stream
                if _pack_stream is None or _pack_stream.device != dev:
                    _pack_stream = named_stream(dev, "pack")
                _pack_stream.wait_stream(torch.cuda.current_stream(dev))
            live.append((k, w, repack))
            self.cache[k] = (w._version, val, ref, repack)
        if dev is None:
            return
        with torch.cuda.stream(_pack_stream):
            single = live
            if BATCH_PACKS:
                batched = [(k, w) for k, w, _ in live if k in self.jobs]
                if batched and self._launch_batch(batched, dev):
                    single = [e for e in live if e[0] not in self.jobs]
            for _, w, repack in single:
                repack(w)
        self._side = _pack_stream

    def _launch_batch(self, batched, dev):
        """One rtp_weight_pack_batch launch for every (key, weight) of `batched`; the device job table is rebuilt only when
        the set of jobs (pointers, shapes) changes — never during a stream capture (then the caller repacks one by one)."""
        import ctypes as C
        sig = tuple((k, w.data_ptr(), self.jobs[k][1].data_ptr()) for k, w in batched)
        if self._batch is None or self._batch[0] != sig:
            if torch.cuda.is_current_stream_capturing():
                return False
            table = (lib.PackJob * len(batched))()
            b0 = 0
            for j, (k, w) in zip(table, batched):
                kind, dst, Cout, Cin_total, ntaps, ci0, ci_n, KP, NP, flag = self.jobs[k]
                total = (ntaps * KP * NP) if kind == 0 else (9 * KP * 3 * NP)
                nb = max(1, min(64, (total + 2047) // 2048))  # >= 8 elements per thread: ~2400 blocks for the whole network
                j.w, j.dst, j.kind, j.Cout, j.Cin_total, j.ntaps = w.data_ptr(), dst.data_ptr(), kind, Cout, Cin_total, ntaps
                j.ci0, j.ci_n, j.KP, j.NP, j.flag, j.block0, j.nblocks = ci0, ci_n, KP, NP, flag, b0, nb
                b0 += nb
            host = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
            self._batch = (sig, host.to(dev), len(batched), b0)
        _, tab, n, nblocks = self._batch
        lib.call("rtp_weight_pack_batch", tab.data_ptr(), n, nblocks, _stream())
        return True

    def invalidate(self):
        """Forces a repack on next use (weights are repacked once per optimizer step in training)."""
        for k in list(self.cache):
            ver, val, ref, repack = self.cache[k]
            self.cache[k] = (None, val, ref, repack)


_pack_stream = None
BATCH_PACKS = not bool(_os.environ.get("RTP_NO_BATCH_PACKS"))  # A/B switch: one launch for all packs of a step


PROFILE = None  # when a dict: key -> list of (start_event, end_event, algorithmic_flops); used by bench.py


def _prof_begin(key):
    if PROFILE is None or (PROFILE.get("_only") and key[0] not in PROFILE["_only"]):
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(key, ev, flops):
    if ev is None:
        return
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    PROFILE.setdefault(key, []).append((ev, end, flops))


def _tbytes(t, C=None):
    """Algorithmic bytes of the real voxels of a P8 tensor (bf16, whole 8-channel chunks are moved)."""
    c8 = ((t.C if C is None else C) + 7) // 8
    return 16.0 * t.N * c8 * t.voxels


# ------------------------------------------------------------------------------------------------ conv
def conv(x, wpack, KP, NP, out, taps, rows, IS=1, OS=1, off=(0, 0, 0), bias=None, res=None, mask=None, relu=False,
         accumulate=False, real=None):
    """Generic tcgen05 implicit-GEMM conv (rtp_conv).  rows = (RZ, RX, RY).  real = (Cin, Cout) for FLOP accounting."""
    d = lib.ConvDesc()
    d.inp, d.out = x.struct(), out.struct()
    d.res = res.struct() if res is not None else lib.NULL_P8
    d.mask = mask.struct() if mask is not None else lib.NULL_P8
    d.w = wpack.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.Cin, d.NP, d.out_c8 = KP, NP, out.C8
    _fill_taps(d, taps)
    d.RZ, d.RX, d.RY = rows
    d.IS, d.OS = IS, OS
    d.oz0, d.ox0, d.oy0 = off
    d.relu, d.accumulate = int(relu), int(accumulate)
    rc = real or (KP, NP)
    key = ("conv_generic", rc[0], rc[1], len(taps), IS, OS, rows)
    ev = _prof_begin(key)
    lib.call("rtp_conv", C.byref(d), _stream())
    _prof_end(key, ev, 2.0 * x.N * rows[0] * rows[1] * rows[2] * rc[0] * rc[1] * len(taps))
    return out


def conv_multi(x, wpack, KP, NP, out, classes, IS=1, OS=1, mask=None, accumulate=False, real=None):
    """Several (taps, rows, offset) classes of one generic conv in ONE launch (rtp_conv_multi): the output-parity classes
    of a stride-2 dgrad."""
    n = len(classes)
    arr = (lib.ConvDesc * n)()
    flops = 0.0
    rc = real or (KP, NP)
    for d, (taps, rows, off) in zip(arr, classes):
        d.inp, d.out = x.struct(), out.struct()
        d.res = lib.NULL_P8
        d.mask = mask.struct() if mask is not None else lib.NULL_P8
        d.w = wpack.data_ptr()
        d.bias = None
        d.Cin, d.NP, d.out_c8 = KP, NP, out.C8
        _fill_taps(d, taps)
        d.RZ, d.RX, d.RY = rows
        d.IS, d.OS = IS, OS
        d.oz0, d.ox0, d.oy0 = off
        d.relu, d.accumulate = 0, int(accumulate)
        flops += 2.0 * x.N * rows[0] * rows[1] * rows[2] * rc[0] * rc[1] * len(taps)
    key = ("conv_generic", rc[0], rc[1], sum(len(c[0]) for c in classes), IS, OS, classes[-1][1])
    ev = _prof_begin(key)
    lib.call("rtp_conv_multi", arr, n, _stream())
    _prof_end(key, ev, flops)
    return out


def pad_bias(b, NP):
    if b is None:
        return None
    if b.numel() == NP:
        return b.detach().float().contiguous()
    o = torch.zeros(NP, dtype=torch.float32, device=b.device)
    o[:b.numel()] = b.detach()
    return o


def out_grid(x, stride):
    return ((x.Z - 1) // stride + 1, (x.Y - 1) // stride + 1, (x.X - 1) // stride + 1) if stride > 1 else x.grid


USE_PW = True    # route 1x1x1 convs through the streaming pointwise kernel (rtp_conv_pw)


def _dense_planes(t):
    return t.c_stride == t.Z * (t.X + 2) * (t.Y + 2) * 8


def pw_eligible(x, out, KP, NP):
    return (x.C8 * 8 >= KP and _dense_planes(x) and _dense_planes(out) and x.grid == out.grid
            and lib.load().rtp_conv_pw_supported(KP, NP) == 1)


def conv_pw(x, wpack, KP, NP, out, bias=None, mask=None, relu=False, accumulate=False, real=None):
    """Pointwise conv / dgrad (rtp_conv_pw)."""
    rc = real or (KP, NP)
    key = ("conv_pw", rc[0], rc[1], 1, 1, 1, (x.Z, x.X, x.Y))
    ev = _prof_begin(key)
    lib.call("rtp_conv_pw", x.struct(), out.struct(), mask.struct() if mask is not None else lib.NULL_P8, wpack.data_ptr(),
             bias.data_ptr() if bias is not None else None, KP, NP, out.C8, int(relu), int(accumulate), _stream())
    _prof_end(key, ev, 2.0 * x.N * x.voxels * rc[0] * rc[1])
    return out


USE_K3S1 = True  # route eligible 3x3x3 stride-1 convs through the plane-streaming kernel (rtp_conv_k3s1)


def k3s1_eligible(x, K, NPo):
    if not USE_K3S1 or x.C8 * 8 != K or x.c_stride != x.Z * (x.X + 2) * (x.Y + 2) * 8:
        return False
    return lib.load().rtp_conv_k3s1_smem_bytes(K, NPo, x.Z, x.X, x.Y) > 0


FUSE_GN_STATS = True  # GroupNorm statistics / backward reductions come out of the producing conv's epilogue


def s2d_tap_mask(par, flipped):
    """In-plane taps (bit kx*3+ky, offsets -1/0/+1 along P8 x / y) that carry weights for parity group `par` of the
    space-to-depth view: parity 0 -> offset 0 only; parity 1 -> offsets -1, 0 (forward) or 0, +1 (dgrad, mirrored)."""
    def valid(bit):
        return (1,) if not bit else ((1, 2) if flipped else (0, 1))
    m = 0
    for tx in valid((par >> 1) & 1):
        for ty in valid(par & 1):
            m |= 1 << (tx * 3 + ty)
    return m


def conv_k3s1(packs, x, w, out, transpose_flip, bias=None, relu=False, res=None, mask=None, accumulate=False, key=None,
              version=None, stat=None, tap_mask=None, ci_window=None, units=None):
    """Plane-streaming 3x3x3 s1 conv (forward: transpose_flip=False; dgrad: True).
    stat: None | ("stats", G, eps) -> returns (out, stats[N][G][2]) of the stored result (GroupNorm forward)
               | ("red", G, x_gn, stats) -> returns (out, red[N][C][2]) (GroupNorm backward reductions, out = dL/dxn)."""
    Cout, Cin = w.shape[0], w.shape[1]
    if ci_window is not None:  # dgrad of a group of input channels: the GEMM N is the window
        Cin = ci_window[1]
    K, NPo = (ceil_to(Cin, 16), ceil_to(Cout, 16)) if not transpose_flip else (ceil_to(Cout, 16), ceil_to(Cin, 16))
    wp = packs.get_k3s1(w, K, NPo, transpose_flip, key, version, ci_window)
    d = lib.ConvK3S1Desc()
    d.inp, d.out = x.struct(), out.struct()
    d.res = res.struct() if res is not None else lib.NULL_P8
    d.mask = mask.struct() if mask is not None else lib.NULL_P8
    d.w = wp.data_ptr()
    b = pad_bias(bias, NPo)
    d.bias = b.data_ptr() if b is not None else None
    d.Cin, d.NPo, d.out_c8 = K, NPo, out.C8
    d.relu, d.accumulate = int(relu), int(accumulate)
    d.debug = None
    d.unit_list, d.unit_count = (units[0].data_ptr(), units[1].data_ptr()) if units is not None else (None, None)
    d.stat_mode, d.stat_ws = 0, None
    d.use_tap_mask = 0
    if tap_mask is not None:  # list of per-K-group in-plane tap masks (structurally sparse weights)
        d.use_tap_mask, d.tap_mask_groups = 1, len(tap_mask)
        for i, m in enumerate(tap_mask):
            d.tap_mask[i] = m
    if stat is not None:
        L = lib.load()
        sws = workspace(L.rtp_conv_k3s1_stat_ws_bytes(x.N), x.buf.device, "k3stat")
        d.stat_mode, d.stat_ws = (1 if stat[0] == "stats" else 2), sws.data_ptr()
        if stat[0] == "red":
            d.stat_aux = stat[2].struct()
    # (launches restricted to a unit list are their own profile family: their work depends on the device-side unit count, so
    # no algorithmic FLOPs are claimed for them and they do not dilute the dense launches' figures)
    pkey = ("conv_k3s1" if units is None else "conv_k3s1_units", Cin if not transpose_flip else Cout,
            Cout if not transpose_flip else Cin, 27, 1, 1, (x.Z, x.X, x.Y))
    ev = _prof_begin(pkey)
    lib.call("rtp_conv_k3s1", C.byref(d), _stream())
    _prof_end(pkey, ev, 2.0 * x.N * x.voxels * Cin * Cout * 27 if units is None else 0.0)
    if stat is None:
        return out
    nct = L.rtp_conv_k3s1_num_ctas(K, NPo, x.N, x.Z, x.X, x.Y)
    Cc, G = out.C, stat[1]
    if stat[0] == "stats":
        res_t = torch.empty((x.N, G, 2), dtype=torch.float32, device=x.buf.device)
        lib.call("rtp_conv_k3s1_stat_finalize", sws.data_ptr(), nct, 1, x.N, Cc, G, out.voxels, float(stat[2]), None,
                 res_t.data_ptr(), _stream())
    else:
        res_t = torch.empty((x.N, Cc, 2), dtype=torch.float32, device=x.buf.device)
        lib.call("rtp_conv_k3s1_stat_finalize", sws.data_ptr(), nct, 2, x.N, Cc, G, out.voxels, 0.0, stat[3].data_ptr(),
                 res_t.data_ptr(), _stream())
    return out, res_t


def stat_fusable(x, w, transpose_flip):
    """True when conv_forward / conv_dgrad of this call goes through the plane-streaming kernel with <= 32 result channels
    (the shapes whose epilogue can carry the GroupNorm reductions)."""
    Cout, Cin = w.shape[0], w.shape[1]
    rc = Cin if transpose_flip else Cout
    kin = Cout if transpose_flip else Cin
    return (FUSE_GN_STATS and w.shape[2] == 3 and rc <= 32 and rc % 8 == 0
            and k3s1_eligible(x, ceil_to(kin, 16), ceil_to(rc, 16)))


def conv_forward(packs, x, w, stride, out, bias=None, relu=False, res=None, ci0=0, ci_n=None, key=None, version=None,
                 stat=None, tap_mask=None, units=None):
    """y = conv3d(x[:, ci0:ci0+ci_n], w[:, ci0:ci0+ci_n], stride, padding=k//2) (+bias)(+res)(relu).
    stat (only when stat_fusable(x, w, False) and stride 1): see conv_k3s1; the return value becomes (y, stats)."""
    k = w.shape[2]
    if k == 3 and stride == 1 and ci0 == 0 and ci_n is None and k3s1_eligible(x, ceil_to(w.shape[1], 16), ceil_to(w.shape[0], 16)):
        return conv_k3s1(packs, x, w, out, False, bias=bias, relu=relu, res=res, key=key, version=version, stat=stat,
                         tap_mask=tap_mask, units=units)
    assert stat is None, "fused statistics need the plane-streaming kernel (check stat_fusable first)"
    assert units is None, "a unit list needs the plane-streaming kernel"
    if k == 1 and stride == 1 and res is None and USE_PW:
        wp, KP, NP = packs.get(w, 0, ci0, ci_n, key, version)
        if pw_eligible(x, out, KP, NP):
            return conv_pw(x, wp, KP, NP, out, bias=pad_bias(bias, NP), relu=relu,
                           real=(ci_n if ci_n is not None else w.shape[1], w.shape[0]))
    wp, KP, NP = packs.get(w, 0, ci0, ci_n, key, version)
    return conv(x, wp, KP, NP, out, taps_fwd(k), (out.Z, out.X, out.Y), IS=stride, bias=pad_bias(bias, NP), res=res,
                relu=relu, real=(ci_n if ci_n is not None else w.shape[1], w.shape[0]))


def conv_dgrad(packs, dy, w, stride, dx, mask=None, accumulate=False, ci0=0, ci_n=None, key=None, version=None, stat=None,
               s2d_cin=None, units=None):
    """dx (=|+=) conv_transpose(dy, w) [* (mask > 0)]; dx has the forward input's geometry.
    stat (only when stat_fusable(dy, w, True) and stride 1): see conv_k3s1; the return value becomes (dx, red)."""
    k = w.shape[2]
    if k == 3 and stride == 1 and ci0 == 0 and ci_n is None and k3s1_eligible(dy, ceil_to(w.shape[0], 16), ceil_to(w.shape[1], 16)):
        return conv_k3s1(packs, dy, w, dx, True, mask=mask, accumulate=accumulate, key=key, version=version, stat=stat, units=units)
    assert stat is None, "fused statistics need the plane-streaming kernel (check stat_fusable first)"
    if (k == 3 and stride == 1 and ci0 == 0 and ci_n is None and w.shape[1] > 80 and w.shape[1] % 32 == 0
            and k3s1_eligible(dy, ceil_to(w.shape[0], 16), 32)):
        # wide dX (e.g. the 128-channel head input, or a space-to-depth view): one plane-streaming launch per group of dX
        # channels.  View groups are paired (64 channels = two parity groups that differ in y: the odd one's taps cover
        # the even one's) when the shape allows, to halve the number of launches.
        gs = 32
        if s2d_cin and w.shape[1] % 64 == 0 and (64 % s2d_cin == 0 or s2d_cin % 64 == 0) and S2D_DGRAD_PAIR \
                and k3s1_eligible(dy, ceil_to(w.shape[0], 16), 64):
            gs = 64
        for g in range(w.shape[1] // gs):
            tm = None
            if s2d_cin:
                tm = 0
                for par in set((c // s2d_cin) for c in range(g * gs, (g + 1) * gs, 8)):
                    tm |= s2d_tap_mask(par, True)
                tm = [tm]
            conv_k3s1(packs, dy, w, dx.channels(g * gs, gs), True,
                      mask=mask.channels(g * gs, gs) if mask is not None else None, accumulate=accumulate,
                      key=(key if key is not None else w.data_ptr(), "dgrad_group", gs, g),
                      version=version if version is not None else w._version, tap_mask=tm, ci_window=(g * gs, gs), units=units)
        return dx
    assert units is None, "a unit list needs the plane-streaming kernel"
    wp, KP, NP = packs.get(w, 1, ci0, ci_n, key, version)
    real = (w.shape[0], ci_n if ci_n is not None else w.shape[1])
    if k == 1 and stride == 1 and USE_PW and pw_eligible(dy, dx, KP, NP):
        return conv_pw(dy, wp, KP, NP, dx, mask=mask, accumulate=accumulate, real=real)
    if stride == 1:
        return conv(dy, wp, KP, NP, dx, taps_dgrad_s1(k), (dx.Z, dx.X, dx.Y), mask=mask, accumulate=accumulate, real=real)
    assert stride == 2 and k == 3
    classes = []
    for pz in range(2):
        for px in range(2):
            for py in range(2):
                rows = ((dx.Z - pz + 1) // 2, (dx.X - px + 1) // 2, (dx.Y - py + 1) // 2)
                if min(rows) > 0:
                    classes.append((taps_dgrad_s2(pz, px, py), rows, (pz, px, py)))
    return conv_multi(dy, wp, KP, NP, dx, classes, IS=1, OS=2, mask=mask, accumulate=accumulate, real=real)


_NSM = None


def num_sms():
    global _NSM
    if _NSM is None:
        _NSM = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _NSM


USE_WGRAD_K3S1 = True
_zero_pages = {}


def _zero_page(nbytes, device):
    z = _zero_pages.get(str(device))
    if z is None or z.numel() < nbytes:
        z = torch.zeros(max(int(nbytes), 16384), dtype=torch.uint8, device=device)
        _zero_pages[str(device)] = z
    return z


def _wgrad_k3s1(x, dy, outs, units=None):
    """Plane-streaming wgrad: loops 32-channel input groups x (<= 48)-channel dY groups; outs = [(dW, acc, ci0, n0)].
    units = (list, count) from active_units(): only those (sample, tile) units are visited (dy is zero in all the others)."""
    L = lib.load()
    dev = x.buf.device
    zero = _zero_page(L.rtp_wgrad_k3s1_zero_bytes(x.Y), dev)
    nsplit = C.c_int32(0)
    for gi in range(x.C // 32):
        xg = x.channels(gi * 32, 32)
        for gw, acc, c0, nn in outs:
            co_total = gw.shape[0]
            h = 0
            while h < co_total:
                hn = (co_total - h) if (co_total - h) <= 48 else 32  # 9 accumulators x NP columns must fit 512
                dyg = dy.channels(nn + h, hn)
                NP = ceil_to(hn, 16)
                ws, done = _split_ws("wgrad3", L.rtp_wgrad_k3s1_workspace_bytes(NP, num_sms()), dev)
                key = ("wgrad_k3s1" if units is None else "wgrad_k3s1_units", 32, hn, 27, 1, 1, (x.Z, x.X, x.Y))
                ev = _prof_begin(key)
                if units is not None:
                    lib.call("rtp_wgrad_k3s1_units", xg.struct(), dyg.struct(), NP, zero.data_ptr(), ws.data_ptr(), C.byref(nsplit),
                             units[0].data_ptr(), units[1].data_ptr(), _stream())
                else:
                    lib.call("rtp_wgrad_k3s1", xg.struct(), dyg.struct(), NP, zero.data_ptr(), ws.data_ptr(), C.byref(nsplit), _stream())
                _prof_end(key, ev, 2.0 * x.N * x.voxels * 32 * hn * 27 if units is None else 0.0)
                done(lambda ws=ws, ns=nsplit.value, NP=NP, gw=gw, h=h, hn=hn, c=c0 + gi * 32, acc=acc: lib.call(
                    "rtp_wgrad_k3s1_reduce", ws.data_ptr(), ns, NP, gw[h:].data_ptr(), gw.shape[1], hn, 0, c, int(acc), _stream()))
                h += hn


import os as _os
_SKIP_WGRAD = bool(_os.environ.get("RTP_SKIP_WGRAD"))
ASYNC_WGRAD = True  # weight gradients run on a side stream (nothing in backward depends on them); see join_wgrad()
_side = {}
LANE = 0  # engines that run concurrently on different streams (half-batch lanes) each get their own weight-gradient stream


WGRAD_STREAM_PER_ORIGIN = not bool(_os.environ.get("RTP_ONE_WGRAD_STREAM"))


def _side_stream(device, kind="wgrad"):
    """Weight-gradient side stream for work forked from the CURRENT stream: one per origin stream (the main stream and each
    branch stream get their own), so the small low-resolution weight gradients of a side branch neither queue behind the
    full-resolution ones nor make them wait for that branch's dY (one shared side stream serialised all of them: with the
    weight gradients skipped the step is 4.8 ms shorter, i.e. they were almost entirely exposed)."""
    origin = _stream() if WGRAD_STREAM_PER_ORIGIN else 0
    key = (str(device), LANE, origin, kind)
    st = _side.get(key)
    if st is None:
        st = {"stream": named_stream(device, "%s%d/%x" % (kind, LANE, origin)), "busy": False, "events": []}
        _side[key] = st
    return st


def join_wgrad(device=None):
    """Makes the current stream wait for every weight gradient queued on the side stream(s)."""
    for key, st in _side.items():
        if st["busy"] and key[1] == LANE and (device is None or key[0] == str(device)):
            cur = torch.cuda.current_stream(st["stream"].device)
            cur.wait_stream(st["stream"])
            if st.get("rstream") is not None:
                cur.wait_stream(st["rstream"])
            st["busy"] = False
            st["ring"] = {}  # everything is joined: no workspace of the ring is still being reduced


def wgrad_streams(device):
    """Every stream weight-gradient work of `device` may be queued on (side streams and their reduction streams)."""
    out = []
    for key, st in _side.items():
        if key[0] == str(device):
            out.append(st["stream"])
            if st.get("rstream") is not None:
                out.append(st["rstream"])
    return out


_cur_side = None  # state of the weight-gradient side stream the running code was forked onto (conv_wgrad_async / on_wgrad_stream)
REDUCE_STREAM = bool(_os.environ.get("RTP_REDUCE_STREAM"))  # opt-in: measured SLOWER (20.52 vs 20.32 ms per step) — the main chain is the critical path


def _split_ws(tag, nbytes, dev):
    """Split-K workspace for a weight-gradient launch plus `done(fn)`, which issues the reduction fn().  On a weight-gradient
    side stream the workspaces alternate between two buffers and the (tiny, latency-bound) reductions go to a companion
    stream, so the next weight-gradient kernel starts right behind the previous one instead of behind its reduction; a
    buffer is reused only after the reduction that read it (event).  Elsewhere: one buffer, reduction in stream order.
    (Experiment, off by default: letting the weight-gradient stream run ahead takes SMs from the main chain and the step
    gets longer, profiles/r02_ab_runs.txt.)"""
    st = _cur_side
    if st is None or not REDUCE_STREAM:
        return workspace(nbytes, dev, tag), (lambda fn: fn())
    ring = st.setdefault("ring", {}).setdefault(tag, {"i": 0, "ev": [None, None]})
    i = ring["i"]
    ring["i"] = i ^ 1
    cur = torch.cuda.current_stream(dev)
    if ring["ev"][i] is not None:
        cur.wait_event(ring["ev"][i])
    ws = workspace(nbytes, dev, "%s#%d" % (tag, i))

    def done(fn):
        if st.get("rstream") is None:
            st["rstream"] = named_stream(dev, "wreduce/%x" % st["stream"].cuda_stream)
        r = st["rstream"]
        r.wait_stream(cur)
        with torch.cuda.stream(r):
            fn()
        ev = torch.cuda.Event()
        ev.record(r)
        ring["ev"][i] = ev
    return ws, done


def conv_wgrad_async(x, dy, k, stride, dW, accumulate=False, ci0=0, n0=0, more=(), then=None, bias_grad=None, units=None):
    """conv_wgrad queued on a side stream, ordered after everything issued so far on the current stream.  The caller
    keeps x, dy and dW alive and untouched until join_wgrad() (the engine's buffers live until the end of the step).
    `then()` is called right after, on the same stream (post-processing of dW).  Falls back to the in-stream call when
    ASYNC_WGRAD is off."""
    dev = x.buf.device
    if _SKIP_WGRAD:  # timing experiments only (RTP_SKIP_WGRAD=1): how much of the step the weight gradients cost net
        return None
    if not ASYNC_WGRAD:
        conv_wgrad(x, dy, k, stride, dW, accumulate, ci0, n0, more, bias_grad=bias_grad, units=units)
        if then is not None:
            then()
        return None
    global _cur_side
    st = _side_stream(dev)
    st["stream"].wait_stream(torch.cuda.current_stream(dev))
    st["busy"] = True
    with torch.cuda.stream(st["stream"]):
        _cur_side = st
        try:
            conv_wgrad(x, dy, k, stride, dW, accumulate, ci0, n0, more, bias_grad=bias_grad, units=units)
            if then is not None:
                then()
        finally:
            _cur_side = None
    return None


def _ones_page(device):
    """128 positions x 8 bf16 channels with channel 0 = 1: the extra GEMM N chunk of rtp_wgrad_pw_bias."""
    z = _ones_pages.get(str(device))
    if z is None:
        z = torch.zeros((128, 8), dtype=torch.bfloat16, device=device)
        z[:, 0] = 1.0
        _ones_pages[str(device)] = z
    return z


_ones_pages = {}
USE_WGRAD_PW_BIAS = not bool(_os.environ.get("RTP_NO_WGRAD_PW_BIAS"))  # A/B switch


def conv_wgrad(x, dy, k, stride, dW, accumulate=False, ci0=0, n0=0, more=(), taps=None, bias_grad=None, units=None):
    """dW[:, ci0:ci0+x.C] (=|+=) wgrad(x, dy[:, n0:n0+dW.shape[0]]).  x: forward input (P8), dy: P8 gradient.
    `more`: further (dW, accumulate, ci0, n0) outputs reduced from the same split-K workspace.
    `taps`: explicit (tz, tx, ty) list instead of the k^3 stencil (the DCN sample volume keeps its taps on the z axis).
    `bias_grad` = (db fp32 [Cout], accumulate): also db (=|+=) sum over positions of dy — inside the streaming 1x1 kernel when it
    takes the shape (a ones channel in the GEMM N), else by a channel_sum pass."""
    if bias_grad is not None and not (USE_WGRAD_PW_BIAS and taps is None and USE_WGRAD_PW and k == 1 and stride == 1 and not more
                                      and n0 == 0 and dy.C <= 128 and dy.C8 <= 16 and x.grid == dy.grid and _dense_planes(x)
                                      and _dense_planes(dy) and dW.shape[0] <= dy.C8 * 8 and dW.shape[0] == bias_grad[0].numel()
                                      and x.C + 8 <= 256 and lib.load().rtp_wgrad_pw_supported(x.C, dy.C, x.Z, x.X, x.Y)):
        channel_sum(dy, bias_grad[0], accumulate=bias_grad[1])
        bias_grad = None
    outs = ((dW, accumulate, ci0, n0),) + tuple(more)
    if (taps is None and USE_WGRAD_K3S1 and k == 3 and stride == 1 and x.C % 32 == 0 and x.c_stride == x.Z * (x.X + 2) * (x.Y + 2) * 8
            and dy.c_stride == x.c_stride and all(o[3] % 8 == 0 for o in outs)
            and lib.load().rtp_wgrad_k3s1_supported(32, 32, x.Z, x.X, x.Y)):
        return _wgrad_k3s1(x, dy, outs, units)
    assert units is None, "a unit list needs the plane-streaming weight-gradient kernel"
    ci_n = x.C
    if (taps is None and USE_WGRAD_PW and k == 1 and stride == 1 and not more and n0 == 0 and dy.C <= 128 and dy.C8 <= 16
            and x.grid == dy.grid and _dense_planes(x) and _dense_planes(dy) and dW.shape[0] <= dy.C8 * 8
            and lib.load().rtp_wgrad_pw_supported(ci_n, dy.C, x.Z, x.X, x.Y)):
        # streaming GEMM over the padded positions (csrc/wgrad_pw.cu): M = dY channels, N = X channels
        L = lib.load()
        dev = x.buf.device
        zero = _zero_page(4096, dev)
        nsplit = C.c_int32(0)
        key = ("wgrad_pw", ci_n, dy.C, 1, 1, 1, (dy.Z, dy.X, dy.Y))
        assert dW.is_contiguous()
        if bias_grad is not None:
            db, accb = bias_grad
            ws, done = _split_ws("wgradpw", L.rtp_wgrad_pw_bias_workspace_bytes(ci_n, num_sms()), dev)
            ev = _prof_begin(key)
            lib.call("rtp_wgrad_pw_bias", x.struct(), dy.struct(), ci_n, zero.data_ptr(), _ones_page(dev).data_ptr(), ws.data_ptr(),
                     C.byref(nsplit), _stream())
            _prof_end(key, ev, 2.0 * dy.N * dy.voxels * ci_n * dy.C)
            done(lambda: lib.call("rtp_wgrad_pw_bias_reduce", ws.data_ptr(), nsplit.value, ci_n, dW.data_ptr(), dW.shape[1], dW.shape[0],
                                  ci0, int(accumulate), db.data_ptr(), int(accb), _stream()))
            return
        ws, done = _split_ws("wgradpw", L.rtp_wgrad_pw_workspace_bytes(ci_n, num_sms()), dev)
        ev = _prof_begin(key)
        lib.call("rtp_wgrad_pw", x.struct(), dy.struct(), ci_n, zero.data_ptr(), ws.data_ptr(), C.byref(nsplit), _stream())
        _prof_end(key, ev, 2.0 * dy.N * dy.voxels * ci_n * dy.C)
        done(lambda: lib.call("rtp_wgrad_pw_reduce", ws.data_ptr(), nsplit.value, ci_n, dW.data_ptr(), dW.shape[1], dW.shape[0], ci0,
                              int(accumulate), _stream()))
        return
    Cin8 = ceil_to(ci_n, 8)
    NP = ceil_to(dy.C, 16)
    taps = taps_fwd(k) if taps is None else taps
    d = lib.WgradDesc()
    d.x, d.dy = x.struct(), dy.struct()
    d.Cin, d.NP = Cin8, NP
    _fill_taps(d, taps, with_wt=False)
    d.RZ, d.RX, d.RY = dy.Z, dy.X, dy.Y
    d.IS = stride
    rows = dy.N * dy.voxels
    ntiles = (rows + 63) // 64
    npairs = len(taps) * (Cin8 // 8)
    nblocks = (npairs + 15) // 16
    per_cta = max(1, 512 // NP)
    groups = (nblocks + per_cta - 1) // per_cta
    nsplit = max(1, min(ntiles, (2 * num_sms()) // groups))
    d.nsplit = nsplit
    ws, done = _split_ws("wgrad", lib.load().rtp_wgrad_workspace_bytes(Cin8, NP, len(taps), nsplit), x.buf.device)
    d.workspace = ws.data_ptr()
    key = ("wgrad_generic", ci_n, dy.C, len(taps), stride, 1, (dy.Z, dy.X, dy.Y))
    ev = _prof_begin(key)
    lib.call("rtp_wgrad", C.byref(d), _stream())
    _prof_end(key, ev, 2.0 * rows * ci_n * dy.C * len(taps))

    def reduce_all():
        for gw, acc, c0, nn in ((dW, accumulate, ci0, n0),) + tuple(more):
            assert gw.is_contiguous()
            lib.call("rtp_wgrad_reduce", ws.data_ptr(), nsplit, Cin8, NP, len(taps), gw.data_ptr(), gw.shape[1], gw.shape[0],
                     nn, c0, ci_n, int(acc), _stream())
    done(reduce_all)


# ------------------------------------------------------------------------------------------------ GroupNorm
def gn_ws(x):
    return workspace(lib.load().rtp_gn_workspace_bytes(x.N, x.C8), x.buf.device, "gn")


def gn_stats(x, G, eps=1e-5):
    stats = torch.empty((x.N, G, 2), dtype=torch.float32, device=x.buf.device)
    if x.C <= 256:
        lib.call("rtp_gn_stats", x.struct(), x.C, G, eps, stats.data_ptr(), gn_ws(x).data_ptr(), _stream())
        return stats
    sums = torch.empty((x.N, x.C, 2), dtype=torch.float32, device=x.buf.device)
    lib.call("rtp_gn_sums", x.struct(), x.C, sums.data_ptr(), gn_ws(x).data_ptr(), _stream())
    lib.call("rtp_gn_finalize", sums.data_ptr(), x.N, x.C, G, x.voxels, eps, stats.data_ptr(), _stream())
    return stats


def gn_apply(x, G, stats, gamma, beta, out):
    key = ("gn_apply", x.C, x.C, 0, 1, 1, (x.Z, x.X, x.Y))
    ev = _prof_begin(key)
    lib.call("rtp_gn_apply", x.struct(), x.C, G, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.struct(),
             _stream())
    _prof_end(key, ev, 2 * _tbytes(x))  # element-wise kernels record algorithmic BYTES in the flops slot
    return out


S2D_MIN_VOXELS = int(_os.environ.get("RTP_S2D_MIN_VOXELS", 1 << 18))
S2D_DGRAD_PAIR = not bool(_os.environ.get("RTP_NO_PAIR"))
USE_WGRAD_PW = not bool(_os.environ.get("RTP_NO_WGRAD_PW"))    # streaming 1x1 weight gradient (csrc/wgrad_pw.cu)
USE_WGRAD_S2D = not bool(_os.environ.get("RTP_NO_WGRAD_S2D"))  # plane-streaming weight gradient over the s2d view (csrc/wgrad_s2d.cu)
USE_S2D = not bool(_os.environ.get("RTP_NO_S2D"))  # stride-2 3x3x3 convs as stride-1 convs over the space-to-depth view (plane-streaming kernels)


def s2d_eligible(x, w):
    """True when conv3d(GN(x), w, stride 2, pad 1) can run through the space-to-depth view: even grid, Cin a multiple of
    8, and the stride-1 problems (forward K = 8*Cin, dgrad in 32-channel groups, wgrad) fit the plane-streaming kernels."""
    Cout, Cin = w.shape[0], w.shape[1]
    if not (USE_S2D and USE_K3S1 and USE_WGRAD_K3S1 and w.shape[2] == 3 and x.C == Cin and Cin % 8 == 0):
        return False
    if x.Z % 2 or x.Y % 2 or x.X % 2 or x.c_stride != x.Z * (x.X + 2) * (x.Y + 2) * 8:
        return False
    if x.N * x.voxels < S2D_MIN_VOXELS:  # small problems are launch-bound: 1 + 8 + 8 stride-1 launches do not pay
        return False
    L = lib.load()
    Z, X, Y = x.Z // 2, x.X // 2, x.Y // 2
    NPo = ceil_to(Cout, 16)
    return (L.rtp_conv_k3s1_smem_bytes(8 * Cin, NPo, Z, X, Y) > 0 and (8 * Cin) % 32 == 0
            and L.rtp_conv_k3s1_smem_bytes(NPo, 32, Z, X, Y) > 0 and L.rtp_wgrad_k3s1_supported(32, 32, Z, X, Y) > 0)


def conv_wgrad_s2d(xs, dy, Cin, dW, accumulate=False):
    """Weight gradient of a stride-2 3x3x3 conv whose (normalised) input is held as the space-to-depth view xs: the 27
    taps of the reference weight are 27 (parity group, offset -1/0) pairs of the view, gathered with unit stride by the
    generic wgrad kernel (per-tap chunk base rtp_wgrad_desc.tc) and written straight into dW[Cout][Cin][3][3][3]."""
    def dim(k):  # tap k of one dimension -> (parity, offset)
        return (1, -1) if k == 0 else ((0, 0) if k == 1 else (1, 0))
    Cin8 = ceil_to(Cin, 8)
    NP = ceil_to(dy.C, 16)
    L = lib.load()
    # the plane-streaming kernel holds 6 accumulators [128 x 2 NP] in TMEM: NP = 32; wider outputs go out as 32-channel slices
    # of dY (dW rows [32 h, 32 h + 32)), each re-reading the L2-resident view
    nslice = NP // 32 if (NP > 32 and dy.C % 32 == 0 and dW.shape[0] == dy.C) else 1
    NPs = 32 if nslice > 1 else NP
    if (USE_WGRAD_S2D and Cin == 32 and xs.C8 == 32 and dy.c_stride == xs.c_stride and dy.C8 * 8 >= NP
            and xs.c_stride == xs.Z * (xs.X + 2) * (xs.Y + 2) * 8 and dW.shape[0] <= NP
            and L.rtp_wgrad_s2d_supported(Cin, NPs, xs.Z, xs.X, xs.Y)):
        # plane-streaming kernel: X staged once per plane, the 27 (parity, offset) pairs are descriptor shifts / N halves
        dev = xs.buf.device
        ws = workspace(L.rtp_wgrad_s2d_workspace_bytes(NPs, num_sms()), dev, "wgrads2d")
        zero = _zero_page(4096, dev)
        assert dW.is_contiguous()
        for h in range(nslice):
            dyh = dy if nslice == 1 else dy.channels(32 * h, 32)
            dWh = dW if nslice == 1 else dW[32 * h:32 * h + 32]
            nsplit = C.c_int32(0)
            key = ("wgrad_s2d", Cin, dyh.C, 27, 2, 1, (dy.Z, dy.X, dy.Y))
            ev = _prof_begin(key)
            lib.call("rtp_wgrad_s2d", xs.struct(), dyh.struct(), Cin, NPs, zero.data_ptr(), ws.data_ptr(), C.byref(nsplit), _stream())
            _prof_end(key, ev, 2.0 * dy.N * dy.voxels * Cin * dyh.C * 27)
            lib.call("rtp_wgrad_s2d_reduce", ws.data_ptr(), nsplit.value, Cin, NPs, dWh.data_ptr(), dWh.shape[0], int(accumulate),
                     _stream())
        return
    d = lib.WgradDesc()
    d.x, d.dy = xs.struct(), dy.struct()
    d.Cin, d.NP = Cin8, NP
    i = 0
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                (pz, oz), (py, oy), (px, ox) = dim(kz), dim(ky), dim(kx)
                d.tz[i], d.tx[i], d.ty[i] = oz, ox, oy
                d.tc[i] = ((pz << 2) | (px << 1) | py) * (Cin8 // 8)
                i += 1
    d.ntaps = 27
    d.RZ, d.RX, d.RY = dy.Z, dy.X, dy.Y
    d.IS = 1
    rows = dy.N * dy.voxels
    ntiles = (rows + 63) // 64
    nblocks = (27 * (Cin8 // 8) + 15) // 16
    per_cta = max(1, 512 // NP)
    groups = (nblocks + per_cta - 1) // per_cta
    nsplit = max(1, min(ntiles, (2 * num_sms()) // groups))
    d.nsplit = nsplit
    ws = workspace(lib.load().rtp_wgrad_workspace_bytes(Cin8, NP, 27, nsplit), xs.buf.device, "wgrad")
    d.workspace = ws.data_ptr()
    key = ("wgrad_generic", Cin, dy.C, 27, 2, 1, (dy.Z, dy.X, dy.Y))
    ev = _prof_begin(key)
    lib.call("rtp_wgrad", C.byref(d), _stream())
    _prof_end(key, ev, 2.0 * rows * Cin * dy.C * 27)
    assert dW.is_contiguous()
    lib.call("rtp_wgrad_reduce", ws.data_ptr(), nsplit, Cin8, NP, 27, dW.data_ptr(), dW.shape[1], dW.shape[0], 0, 0, Cin,
             int(accumulate), _stream())


def on_wgrad_stream(x, fn):
    """Runs fn() on the weight-gradient side stream (ordered after the current stream), or in place when ASYNC_WGRAD is off."""
    dev = x.buf.device
    if _SKIP_WGRAD:
        return
    if not ASYNC_WGRAD:
        return fn()
    global _cur_side
    st = _side_stream(dev)
    st["stream"].wait_stream(torch.cuda.current_stream(dev))
    st["busy"] = True
    with torch.cuda.stream(st["stream"]):
        _cur_side = st
        try:
            fn()
        finally:
            _cur_side = None


def on_aux_stream(x, fn):
    """Runs fn() on the auxiliary parameter-gradient stream (ordered after the current stream; joined by join_wgrad): the
    HBM-bound bias-gradient channel sums, which nothing in backward depends on, then run beside the tensor-bound
    weight-gradient kernels instead of in front of the next dgrad of the main chain."""
    dev = x.buf.device
    if not ASYNC_WGRAD or not AUX_STREAM:
        return fn()
    st = _side_stream(dev, "aux")
    st["stream"].wait_stream(torch.cuda.current_stream(dev))
    st["busy"] = True
    with torch.cuda.stream(st["stream"]):
        fn()


AUX_STREAM = bool(_os.environ.get("RTP_AUX_STREAM"))  # opt-in: measured slightly SLOWER (19.85 vs 19.75 ms per step) — the sums stretch the weight-gradient kernels they run beside


def gn_apply_s2d(x, G, stats, gamma, beta, out):
    """GroupNorm apply that writes the result as the space-to-depth view `out` (grid halved, 8x the channels)."""
    lib.call("rtp_gn_apply_s2d", x.struct(), x.C, G, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.struct(),
             _stream())
    return out


def s2d_expand(w, out=None):
    """[Cout][Cin][3][3][3] -> the weights [Cout][8*Cin][3][3][3] of the equivalent stride-1 conv over the s2d view."""
    Cout, Cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty((Cout, 8 * Cin, 3, 3, 3), dtype=torch.float32, device=w.device)
    lib.call("rtp_weight_s2d_expand", w.detach().contiguous().data_ptr(), out.data_ptr(), Cout, Cin, _stream())
    return out


def s2d_fold(dw_s2d, dw, accumulate):
    lib.call("rtp_weight_s2d_fold", dw_s2d.data_ptr(), dw.data_ptr(), dw.shape[0], dw.shape[1], int(accumulate), _stream())


USE_S2D_SHARE = not bool(_os.environ.get("RTP_NO_S2D_SHARE"))  # sibling stride-2 convs share one view of xhat (csrc/s2d_shared.cu)


def s2d_fold_weights(w, gamma, beta, we, bias_cls):
    """we [Cout][8*Cin][27] = s2d_expand(W * diag(gamma)); bias_cls [8][Cout] = the beta term per border class."""
    lib.call("rtp_s2d_fold_weights", w.detach().contiguous().data_ptr(), gamma.data_ptr(), beta.data_ptr(), we.data_ptr(),
             bias_cls.data_ptr(), w.shape[0], w.shape[1], _stream())


def s2d_border_bias(bias_cls, r, Cout):
    """r (P8, N = 1) = bias_cls[class(pos)] - bias_cls[0]; returns r broadcast over the batch (n_stride = 0) for use as the
    conv's `res` input."""
    lib.call("rtp_s2d_border_bias", bias_cls.data_ptr(), r.struct(), Cout, _stream())


def s2d_fold_wgrad(dy, dw_xhat, w, gamma, beta, dW, dgamma, dbeta, acc_w, acc_gb):
    """(dW, dgamma, dbeta) of a gamma/beta-folded stride-2 conv from the weight gradient over xhat (csrc/s2d_shared.cu)."""
    Cout, Cin = w.shape[0], w.shape[1]
    ws = workspace(lib.load().rtp_s2d_box_sums_workspace_bytes(dy.N, (Cout + 7) // 8), dy.buf.device, "s2dbox")
    assert dW.is_contiguous() and dw_xhat.is_contiguous()
    lib.call("rtp_s2d_fold_wgrad", dy.struct(), dw_xhat.data_ptr(), w.detach().contiguous().data_ptr(), gamma.data_ptr(),
             beta.data_ptr(), ws.data_ptr(), dW.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), Cout, Cin, int(acc_w), int(acc_gb),
             _stream())


def gn_backward(x, dxn, G, stats, gamma, dgamma, dbeta, acc_params, dx, acc_dx, red=None, s2d=False, add=None):
    """red: the [N][C][2] reductions when the conv that produced dxn already computed them (conv_dgrad(stat=...)).
    s2d: dxn is the gradient of the space-to-depth view of GN(x).
    add: a second gradient into x (same geometry; masked with (x > 0) like the GroupNorm term when x is a ReLU output)."""
    sfx = "_s2d" if s2d else ""
    key = ("gn_backward", x.C, x.C, 0 if red is None else 1, 1, 1, (x.Z, x.X, x.Y))
    ev = _prof_begin(key)
    if red is None:
        red = torch.empty((x.N, x.C, 2), dtype=torch.float32, device=x.buf.device)
        lib.call("rtp_gn_bwd_reduce" + sfx, x.struct(), dxn.struct(), x.C, G, stats.data_ptr(), red.data_ptr(),
                 gn_ws(x).data_ptr(), _stream())
    lib.call("rtp_gn_bwd_apply" + sfx, x.struct(), dxn.struct(), x.C, G, stats.data_ptr(), red.data_ptr(), gamma.data_ptr(),
             dgamma.data_ptr() if dgamma is not None else None, dbeta.data_ptr() if dbeta is not None else None,
             int(acc_params), dx.struct() if dx is not None else lib.NULL_P8,
             int(acc_dx), int(x.relu_out), add.struct() if add is not None else lib.NULL_P8, _stream())
    # bytes: (reduction pass: x + dxn) + (apply pass: x + dxn read, dx written [+ read when accumulating] [+ add read])
    nb = (0 if key[3] else 2) + (0 if dx is None else 3 + (1 if acc_dx else 0) + (1 if add is not None else 0))
    _prof_end(key, ev, nb * _tbytes(x))


# ------------------------------------------------------------------------------------------------ fuse / misc
def fuse_sum(out, same, low, bias=None, relu=False):
    d = lib.FuseDesc()
    d.out, d.C = out.struct(), out.C
    d.n_same, d.n_low = len(same), len(low)
    for i, t in enumerate(same):
        d.same[i] = t.struct()
    for i, t in enumerate(low):
        d.low[i] = t.struct()
    d.bias = bias.data_ptr() if bias is not None else None
    d.relu = int(relu)
    key = ("fuse_sum", out.C, out.C, len(same) + len(low), 1, 1, (out.Z, out.X, out.Y))
    ev = _prof_begin(key)
    lib.call("rtp_fuse_sum", C.byref(d), _stream())
    _prof_end(key, ev, _tbytes(out) * (1 + len(same)) + sum(_tbytes(t, out.C) for t in low))
    return out


USE_CONAT = not bool(_os.environ.get("RTP_NO_CONAT"))  # final concat + 1x1 conv as one GEMM with an interpolated A operand (csrc/conat.cu)


def conat_forward(packs, ys, w, bias, out, relu=False):
    """out = [relu](conv1x1(cat(ys[0], up(ys[1]), ...), w) + bias) in ONE launch (rtp_conat_fwd), or None when the shape is
    not supported (the caller then uses per-branch 1x1 convs + fuse_sum)."""
    if not USE_CONAT or not (2 <= len(ys) <= 4) or w.shape[2] != 1:
        return None
    Cout, Cin = w.shape[0], w.shape[1]
    if sum(y.C for y in ys) != Cin or any(y.C % 16 for y in ys) or Cout % 16:
        return None
    x0 = ys[0]
    if not (_dense_planes(x0) and _dense_planes(out) and x0.grid == out.grid):
        return None
    d = lib.ConatDesc()
    d.x0, d.out = x0.struct(), out.struct()
    d.n_low, d.c_x0 = len(ys) - 1, x0.C
    for j, y in enumerate(ys[1:]):
        d.low[j] = y.struct()
        d.c_low[j] = y.C
    d.K, d.NP, d.out_c8, d.relu = Cin, Cout, Cout // 8, int(relu)
    if lib.load().rtp_conat_supported(C.byref(d)) != 1:
        return None
    wp, KP, NP = packs.get(w, 0)
    assert (KP, NP) == (Cin, Cout)
    d.w = wp.data_ptr()
    b = pad_bias(bias, NP)
    d.bias = b.data_ptr() if b is not None else None
    key = ("conat", Cin, Cout, 1, 1, 1, (out.Z, out.X, out.Y))
    ev = _prof_begin(key)
    lib.call("rtp_conat_fwd", C.byref(d), _stream())
    _prof_end(key, ev, _tbytes(x0) + _tbytes(out) + sum(_tbytes(y) for y in ys[1:]))  # algorithmic BYTES (HBM-bound)
    return out


def upsample_bwd(dout, dlow, accumulate=False):
    ws = workspace(lib.load().rtp_upsample_bwd_workspace_bytes(dout.struct(), dlow.struct(), dlow.C), dout.buf.device, "upbwd")
    key = ("upsample_bwd", dlow.C, dlow.C, 0, 1, 1, (dout.Z, dout.X, dout.Y))
    ev = _prof_begin(key)
    lib.call("rtp_upsample_bwd", dout.struct(), dlow.struct(), dlow.C, int(accumulate), ws.data_ptr(), _stream())
    _prof_end(key, ev, _tbytes(dout, dlow.C) + _tbytes(dlow) * (2 if accumulate else 1))


def grad_add(src, dst, mask=None, accumulate=False):
    key = ("grad_add", dst.C, dst.C, 0, 1, 1, (dst.Z, dst.X, dst.Y))
    ev = _prof_begin(key)
    lib.call("rtp_grad_add", src.struct(), mask.struct() if mask is not None else lib.NULL_P8, dst.struct(), dst.C,
             int(accumulate), _stream())
    _prof_end(key, ev, _tbytes(dst) * (2 + (1 if mask is not None else 0) + (1 if accumulate else 0)))


USE_SPARSE_REG = not bool(_os.environ.get("RTP_NO_SPARSE_REG"))  # A/B switch: rtp_reg_head_bwd_sparse
USE_SPARSE_UNITS = not bool(_os.environ.get("RTP_NO_SPARSE_UNITS"))  # A/B switch: unit lists for the regression half of the head
USE_SPARSE_DREG = not bool(_os.environ.get("RTP_NO_SPARSE_DREG"))  # A/B switch: the loss writes d_reg at the target voxels only
USE_SPARSE_FWD = not bool(_os.environ.get("RTP_NO_SPARSE_FWD"))  # A/B switch: training forward of the regression branch on the listed units only


def active_units(ind, like, radius, tag):
    """(unit list int32 [N * ntile], count int32 [1]) of the (sample, 128-position tile) units of `like`'s grid that contain a
    voxel within `radius` (x, y) of a target voxel ind[n][m] — the units conv_k3s1 / wgrad_k3s1 have to visit when their input
    is zero everywhere else (rtp_active_units).  The buffers belong to (tag, stream) and are reused every step."""
    N, M = ind.shape
    ntile = (like.X * (like.Y + 2) + 127) // 128
    buf = workspace(4 * (N * ntile + 4), like.buf.device, "units_" + tag).view(torch.int32)
    lst, cnt = buf[:N * ntile], buf[N * ntile:N * ntile + 1]
    lib.call("rtp_active_units", ind.data_ptr(), N, M, like.Z, like.X, like.Y, radius, lst.data_ptr(), cnt.data_ptr(), _stream())
    return lst, cnt


USE_PREZERO = not bool(_os.environ.get("RTP_NO_PREZERO"))  # A/B switch: zero-fill of the hidden regression gradient beside the head convs


def prezero(x):
    """Zero-fills the chunk volumes of x on a side stream forked from the current one (rtp_zero_chunks) and returns the event
    to wait for before x is used (None when the fill was issued in line).  Engine.head uses it for the regression half of the
    head's hidden gradient: an HBM-bound fill that otherwise sits between the loss and the first backward kernel now runs
    beside the tensor-bound head convolutions of the forward pass."""
    dev = x.buf.device
    if not ASYNC_WGRAD:
        lib.call("rtp_zero_chunks", x.struct(), _stream())
        return None
    st = _side_stream(dev, "zero")
    st["stream"].wait_stream(torch.cuda.current_stream(dev))
    st["busy"] = True
    with torch.cuda.stream(st["stream"]):
        lib.call("rtp_zero_chunks", x.struct(), _stream())
        ev = torch.cuda.Event()
        ev.record(st["stream"])
    return ev


def reg_head_bwd_sparse(d_reg, t_in, ind, w, dt, dW, acc_w, db, acc_b, prezeroed=False):
    """Backward of the regression branch's last 3x3x3 conv from the sparse loss gradient (csrc/head_sparse.cu).
    prezeroed: dt's chunks are zero already (prezero above)."""
    L = lib.load()
    N, M = ind.shape
    R, Cin = w.shape[0], w.shape[1]
    ws = workspace(L.rtp_reg_head_bwd_sparse_workspace_bytes(N, M), d_reg.buf.device, "regsp")
    wc = w.detach()
    assert wc.is_contiguous() and dW.is_contiguous() and ind.dtype == torch.int64 and ind.is_contiguous()
    lib.call("rtp_reg_head_bwd_sparse_prezeroed" if prezeroed else "rtp_reg_head_bwd_sparse", d_reg.struct(), t_in.struct(),
             ind.data_ptr(), M, wc.data_ptr(), R, Cin, dt.struct(),
             dW.data_ptr(), int(acc_w), db.data_ptr(), int(acc_b), ws.data_ptr(), _stream())


def reg_sparse_supported(w, ind):
    return (USE_SPARSE_REG and w.dim() == 5 and tuple(w.shape[2:]) == (3, 3, 3) and w.shape[0] <= 64 and ind is not None
            and ind.dim() == 2 and ind.shape[1] <= 64 and ind.dtype == torch.int64)


def channel_sum(x, out, accumulate=False):
    lib.call("rtp_channel_sum", x.struct(), x.C, out.data_ptr(), int(accumulate), gn_ws(x).data_ptr(), _stream())
