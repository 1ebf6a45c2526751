"""HRRadarPose forward/backward as a straight-line program of C-ABI kernel launches.

This is the host side of the hot path: it mirrors the dataflow of the reference modules
(det3d/models/backbones/hr_util/{common,hr3d}.py, backbones/hrnet3d.py, pose_heads/center_head.py) but runs
every op through librtpose_b200.so on P8 tensors and keeps its own tape for the backward pass (the graph is
static, so no autograd engine is needed; `autograd_bridge.py` exposes it to torch as one Function).

Gradient convention: for a tensor that is the output of a ReLU, `.grad` holds the gradient w.r.t. the *pre-ReLU*
value; every kernel that writes into such a gradient applies the (t > 0) mask itself.
"""
import contextlib
import os

import torch

from . import lib, ops
from .p8 import P8, _stream

ARCH = {  # det3d/models/backbones/hrnet3D_config.py:85-177 — (input planes, per-branch channels)
    "hr_tiny_feat32_zyx_l4": (1, [32, 32, 64, 64]),
    "hr_tiny_feat32_zyx_l4_in32": (32, [32, 32, 64, 64]),
    "hr_tiny_feat64_zyx_l4_in64": (64, [64, 64, 128, 128]),
}


def _adjacent_cat(a, b):
    """cat([a, b], 0) of two parameters — as a VIEW when b starts where a ends in the same storage (flat-buffer trainers,
    bench.build_params places the two merged head convs next to each other), else a copy (torch.cat)."""
    a, b = a.detach(), b.detach()
    if (a.is_contiguous() and b.is_contiguous() and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
            and a.storage_offset() + a.numel() == b.storage_offset() and a.shape[1:] == b.shape[1:]):
        return torch.as_strided(a, (a.shape[0] + b.shape[0],) + tuple(a.shape[1:]), a.stride())
    return torch.cat([a, b], 0).contiguous()


FUSE_PRIORITY = int(os.environ.get("RTP_FUSE_PRIO", "-2"))  # stream priority of the side fuse outputs (0: the branch streams, as before)
BRANCH_PRIORITY = int(os.environ.get("RTP_BRANCH_PRIO", "-1"))  # stream priority of the side branches b >= 1 (res blocks): the capture stream's, not below it
FUSE_LONGEST_FIRST = not bool(os.environ.get("RTP_NO_FUSE_ORDER"))  # A/B switch: issue the side fuse outputs longest chain first


class Engine:
    def __init__(self, arch, final_fuse, params, reg_channels, num_classes, loss_weight, code_weights,
                 prefix_backbone="backbone.", prefix_head="pose_head."):
        """params: dict name -> fp32 CUDA tensor (the reference's state_dict names)."""
        if arch not in ARCH:
            raise KeyError("unknown backbone_cfg %r (supported: %s)" % (arch, sorted(ARCH)))
        self.in_ch, self.ch = ARCH[arch]
        if torch.cuda.is_available():
            lib.setup_device()  # device-wide shared-memory preference (lib.setup_device)
        self.fuse = final_fuse
        self.p = params
        self.pb, self.ph = prefix_backbone, prefix_head
        self.R, self.ncls = reg_channels, num_classes
        self.loss_weight = float(loss_weight)
        self.code_weights = [float(v) for v in code_weights]
        self.packs = ops.PackedWeights()
        self.pool = ops.BufferPool()
        self.tape = []
        self.grads = None
        self._touched = set()
        self._cw = None
        # the branches of an HR module are independent between two fuse layers: run branch b >= 1 on its own stream
        # so the small low-resolution kernels fill the gaps of the full-resolution branch (forward and backward)
        self.parallel_branches = True
        self.parallel_fuse = True
        self.parallel_fuse_bwd = False  # measured: no gain (the ordered accumulation into shared gradients serialises it)
        self._ordered_grads = False  # set while backward closures may run on several streams
        self._bstreams = {}
        self.fold_grad_adds = not os.environ.get("RTP_NO_FOLD_ADDS")  # residual / fuse-sum gradient pass-throughs ride in the next GroupNorm backward (_defer_add)
        self._last_touch, self._closure_idx = {}, -1
        self._grad_groups, self._on_ready, self._milestones = None, None, None
        self._s2d_share, self._s2d_scratch, self._unit_affine = {}, {}, {}
        self._reg_sparse = None  # (reg.grad, ind, sparse_only) of the last loss() call: lets head.bwd take the sparse path
        self._reg_out = self._reg2_weight = None  # the training head's regression output and last conv weight (head -> loss)
        self.generation = 0  # bumped by begin(): a backward job checks that the tape it recorded is still the live one

    # ------------------------------------------------------------------ helpers
    def new(self, like, C=None, grid=None):
        Z, Y, X = grid if grid is not None else like.grid
        return self.pool.get(like.N, like.C if C is None else C, Z, Y, X, like.buf.device)

    # GroupNorm statistics are computed once per tensor and shared by every GN that reads it.  The cache is keyed by the
    # tensor OBJECT and keeps it alive: keyed by id() alone, a temporary P8 that died (a channel view, a converted input)
    # could hand its statistics to a new tensor allocated at the same address.
    def _stats_get(self, t):
        hit = self.stats_cache.get(id(t))
        if hit is not None and hit[0] is t and hit[2] == (t.buf.data_ptr(), t.offset):
            return hit[1]
        return None

    def _stats_put(self, t, stats):
        self.stats_cache[id(t)] = (t, stats, (t.buf.data_ptr(), t.offset))

    def _grad_of(self, t):
        """Returns (grad tensor, accumulate flag) for a writer into t.grad.  Writers on different streams are chained in
        program order: the current stream first waits for the previous writer's event; call _wrote(t) after the write."""
        if t.grad_ev is not None:
            torch.cuda.current_stream(t.buf.device).wait_event(t.grad_ev)
        if t.grad is None:
            t.grad = self.new(t)
            return t.grad, False
        return t.grad, True

    def _wrote(self, t):
        if self._ordered_grads:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.buf.device))
            t.grad_ev = ev

    def _defer_add(self, t, src):
        """t.grad += src (times (t > 0) when t is a ReLU output) — the pass-through of a residual / fuse sum — WITHOUT a pass
        of its own: the next GroupNorm backward that writes t.grad takes `src` as its `add` input (rtp_gn_bwd_apply).  If
        t.grad is read before such a writer comes, _g() performs the add with rtp_grad_add."""
        if not self.fold_grad_adds:
            gt, acc = self._grad_of(t)
            ops.grad_add(src, gt, mask=t if t.relu_out else None, accumulate=acc)
            self._wrote(t)
            return
        self._flush_pending(t)
        t.pending = src

    def _flush_pending(self, t):
        if t.pending is not None:
            src, t.pending = t.pending, None
            gt, acc = self._grad_of(t)
            ops.grad_add(src, gt, mask=t if t.relu_out else None, accumulate=acc)
            self._wrote(t)

    def _g(self, t):
        """t.grad for a READER (every deferred add applied first)."""
        self._flush_pending(t)
        if t.grad_ev is not None:  # written on another stream (stream-parallel fuse backward)
            torch.cuda.current_stream(t.buf.device).wait_event(t.grad_ev)
        return t.grad

    def _take_pending(self, t):
        src, t.pending = t.pending, None
        return src

    def _pgrad(self, name):
        """fp32 gradient tensor of parameter `name` and whether to accumulate into it."""
        g = self.grads[name]
        acc = name in self._touched
        self._touched.add(name)
        self._last_touch[name] = self._closure_idx
        return g, acc

    def set_grad_groups(self, groups, on_ready):
        """Gradient-ready notifications for overlapping the data-parallel all-reduce with the backward pass
        (what DDP's buckets do for the reference, det3d/torchie/apis/train.py:285-291).  `groups`: lists of parameter
        names; `on_ready(k)` is called inside backward() right after the LAST tape closure that writes a gradient of group
        k has been issued (its kernels are queued on the current / weight-gradient / branch streams — the callback
        makes its communication stream wait for them, see dist.SlicedAllReduce).  The closure -> group map is learnt from
        the previous backward pass (the tape is static); until then every group is reported at the end of backward()."""
        self._grad_groups = [list(g) for g in groups]
        self._on_ready = on_ready
        self._milestones = None

    # ------------------------------------------------------------------ ops with backward closures
    def gn_conv(self, x, gn, conv, k, stride, relu, res=None, train=True, x_needs_grad=True, res_needs_grad=True,
                y_stats=False):
        """ReLU?( conv_k_stride( GroupNorm(8)(x) ) [+ res] ) — common.py:92-96 'gcr'/'gc' order; hr3d.py fuse/transition
        Sequentials.  y_stats: the result feeds another GroupNorm, so let the conv's epilogue produce its statistics
        when the shape allows (ops.stat_fusable); likewise the dgrad's epilogue produces the GroupNorm-backward sums."""
        p = self.p
        gamma, beta, w = p[gn + ".weight"], p[gn + ".bias"], p[conv + ".weight"]
        G = 8 if x.C >= 8 else 1
        stats = self._stats_get(x)
        if stats is None:
            stats = ops.gn_stats(x, G)
            self._stats_put(x, stats)
        Zo, Yo, Xo = ops.out_grid(x, stride)
        y = self.new(x, C=w.shape[0], grid=(Zo, Yo, Xo))
        if stride == 2 and k == 3 and ops.s2d_eligible(x, w):
            return self._gn_conv_s2d(x, y, G, stats, gamma, beta, w, gn, conv, relu, res, train, x_needs_grad, res_needs_grad)
        xn = ops.gn_apply(x, G, stats, gamma, beta, self.new(x))
        if y_stats and stride == 1 and y.C % 8 == 0 and ops.stat_fusable(xn, w, False):
            _, st = ops.conv_forward(self.packs, xn, w, stride, y, relu=relu, res=res, stat=("stats", 8, 1e-5))
            self._stats_put(y, st)
        else:
            ops.conv_forward(self.packs, xn, w, stride, y, relu=relu, res=res)
        y.relu_out = bool(relu)
        if train:
            def bwd():
                dy = self._g(y)
                if dy is None:
                    return
                if res is not None and res_needs_grad:
                    self._defer_add(res, dy)
                gw, accw = self._pgrad(conv + ".weight")
                ops.conv_wgrad_async(xn, dy, k, stride, gw, accumulate=accw)
                red = None
                if stride == 1 and x.C % 8 == 0 and ops.stat_fusable(dy, w, True):
                    dxn, red = ops.conv_dgrad(self.packs, dy, w, stride, self.new(xn), stat=("red", G, x, stats))
                else:
                    dxn = ops.conv_dgrad(self.packs, dy, w, stride, self.new(xn))
                gg, accg = self._pgrad(gn + ".weight")
                gb, _ = self._pgrad(gn + ".bias")
                if x_needs_grad:
                    gx, accx = self._grad_of(x)
                else:
                    gx, accx = None, False
                ops.gn_backward(x, dxn, G, stats, gamma, gg, gb, accg, gx, accx, red=red,
                                add=self._take_pending(x) if gx is not None else None)
                if gx is not None:
                    self._wrote(x)
            self.tape.append(bwd)
        return y

    def _gn_conv_s2d(self, x, y, G, stats, gamma, beta, w, gn, conv, relu, res, train, x_needs_grad, res_needs_grad):
        """Stride-2 3x3x3 gn_conv through the space-to-depth view (include/rtpose_b200.h, rtp_gn_apply_s2d): GroupNorm
        writes the view, the conv / dgrad / wgrad are stride-1 plane-streaming launches with the expanded weights, and
        GroupNorm backward reads the gradient of the view."""
        dev = x.buf.device
        xs = self.pool.get(x.N, 8 * x.C, x.Z // 2, x.Y // 2, x.X // 2, dev)
        ops.gn_apply_s2d(x, G, stats, gamma, beta, xs)
        we = ops.s2d_expand(w)
        wkey, wver = ("s2d", w.data_ptr()), w._version
        ops.conv_forward(self.packs, xs, we, 1, y, relu=relu, res=res, key=wkey, version=wver,
                         tap_mask=[ops.s2d_tap_mask(par, False) for par in range(8)])
        y.relu_out = bool(relu)
        if train:
            def bwd():
                dy = self._g(y)
                if dy is None:
                    return
                if res is not None and res_needs_grad:
                    self._defer_add(res, dy)
                gw, accw = self._pgrad(conv + ".weight")
                ops.on_wgrad_stream(xs, lambda: ops.conv_wgrad_s2d(xs, dy, x.C, gw, accumulate=accw))
                dxs = ops.conv_dgrad(self.packs, dy, we, 1, self.pool.get(x.N, 8 * x.C, x.Z // 2, x.Y // 2, x.X // 2, dev),
                                     key=wkey, version=wver, s2d_cin=x.C)
                gg, accg = self._pgrad(gn + ".weight")
                gb, _ = self._pgrad(gn + ".bias")
                if x_needs_grad:
                    gx, accx = self._grad_of(x)
                else:
                    gx, accx = None, False
                ops.gn_backward(x, dxs, G, stats, gamma, gg, gb, accg, gx, accx, s2d=True,
                                add=self._take_pending(x) if gx is not None else None)
                if gx is not None:
                    self._wrote(x)
            self.tape.append(bwd)
        return y

    def _s2d_share_prepare(self, x, n_sib):
        """One space-to-depth view of xhat = (x - mean) * rstd for the `n_sib` stride-2 fuse convs that read x (their
        GroupNorms share x's statistics; the affine of each is folded into its conv, csrc/s2d_shared.cu)."""
        dev = x.buf.device
        G = 8 if x.C >= 8 else 1
        stats = self._stats_get(x)
        if stats is None:
            stats = ops.gn_stats(x, G)
            self._stats_put(x, stats)
        key = (x.C, str(dev))
        if key not in self._unit_affine:
            self._unit_affine[key] = (torch.ones(x.C, dtype=torch.float32, device=dev),
                                      torch.zeros(x.C, dtype=torch.float32, device=dev))
        ones, zeros = self._unit_affine[key]
        V = self.pool.get(x.N, 8 * x.C, x.Z // 2, x.Y // 2, x.X // 2, dev)
        ops.gn_apply_s2d(x, G, stats, ones, zeros, V)
        return {"x": x, "V": V, "G": G, "stats": stats, "ones": ones, "dV": None, "left": n_sib}

    def _gn_conv_s2d_shared(self, grp, gn, conv, relu, train):
        """ReLU?( conv3x3x3_stride2( GroupNorm(x) ) ) for one of the sibling fuse convs of grp["x"], reading the shared view:
        y = conv_{W diag(gamma)}(xhat view) + border-class bias (the beta term).  Backward: the weight gradient over xhat gives
        dW / dgamma / dbeta (rtp_s2d_fold_wgrad), the dgrads of the siblings accumulate into one dL/dxhat view, and the last
        sibling to run its closure issues the single GroupNorm backward into x."""
        p = self.p
        x, V = grp["x"], grp["V"]
        dev = x.buf.device
        gamma, beta, w = p[gn + ".weight"], p[gn + ".bias"], p[conv + ".weight"]
        Cout, Cin = w.shape[0], w.shape[1]
        sc = self._s2d_scratch.get((conv, str(dev)))
        if sc is None:
            sc = (torch.empty((Cout, 8 * Cin, 3, 3, 3), dtype=torch.float32, device=dev),
                  torch.empty((8, Cout), dtype=torch.float32, device=dev),
                  torch.empty((Cout, Cin, 3, 3, 3), dtype=torch.float32, device=dev))
            self._s2d_scratch[(conv, str(dev))] = sc
        we, bias_cls, dwp = sc
        ops.s2d_fold_weights(w, gamma, beta, we, bias_cls)
        y = self.pool.get(x.N, Cout, V.Z, V.Y, V.X, dev)
        r1 = self.pool.get(1, Cout, V.Z, V.Y, V.X, dev)
        ops.s2d_border_bias(bias_cls, r1, Cout)
        rb = P8(x.N, Cout, V.Z, V.Y, V.X, buf=r1.buf, offset=r1.offset, n_stride=0, c_stride=r1.c_stride)
        wkey, wver = ("s2dgn", w.data_ptr()), (w._version, gamma._version, beta._version)
        ops.conv_forward(self.packs, V, we, 1, y, bias=bias_cls[0], relu=relu, res=rb, key=wkey, version=wver,
                         tap_mask=[ops.s2d_tap_mask(par, False) for par in range(8)])
        y.relu_out = bool(relu)
        if train:
            def bwd():
                grp["left"] -= 1
                dy = self._g(y)
                if dy is not None:
                    gw, accw = self._pgrad(conv + ".weight")
                    gg, accg = self._pgrad(gn + ".weight")
                    gb, _ = self._pgrad(gn + ".bias")

                    def wg():
                        ops.conv_wgrad_s2d(V, dy, Cin, dwp, accumulate=False)
                        ops.s2d_fold_wgrad(dy, dwp, w, gamma, beta, gw, gg, gb, accw, accg)
                    ops.on_wgrad_stream(V, wg)
                    first = grp["dV"] is None
                    if first:
                        grp["dV"] = self.pool.get(x.N, 8 * Cin, V.Z, V.Y, V.X, dev)
                    ops.conv_dgrad(self.packs, dy, we, 1, grp["dV"], accumulate=not first, key=wkey, version=wver, s2d_cin=Cin)
                if grp["left"] == 0 and grp["dV"] is not None:
                    gx, accx = self._grad_of(x)
                    ops.gn_backward(x, grp["dV"], grp["G"], grp["stats"], grp["ones"], None, None, False, gx, accx, s2d=True,
                                    add=self._take_pending(x))
                    self._wrote(x)
            self.tape.append(bwd)
        return y

    def res_block(self, x, prefix, train, x_needs_grad=True):
        """ResNetBlock.forward (hr_util/common.py:138-148)."""
        p = self.p
        if prefix + ".conv1.weight" in p:
            w1, b1 = p[prefix + ".conv1.weight"], p[prefix + ".conv1.bias"]
            if w1.shape[1] != 1:
                raise NotImplementedError("ResNetBlock.conv1 with Cin=%d (only the 1->C stem occurs)" % w1.shape[1])
            r = self.new(x, C=w1.shape[0])
            wv = w1.detach().reshape(-1).contiguous()
            lib.call("rtp_stem_fwd", x.struct(), wv.data_ptr(), b1.data_ptr(), w1.shape[0], r.struct(), _stream())
            if train:
                def bwd():
                    if self._g(r) is None:
                        return
                    gw, acc = self._pgrad(prefix + ".conv1.weight")
                    gb, _ = self._pgrad(prefix + ".conv1.bias")
                    lib.call("rtp_stem_bwd", x.struct(), r.grad.struct(), w1.shape[0], gw.data_ptr(), gb.data_ptr(),
                             int(acc), ops.gn_ws(r).data_ptr(), _stream())
                self.tape.append(bwd)
            r_needs_grad = True
        else:
            r, r_needs_grad = x, x_needs_grad
        o = self.gn_conv(r, prefix + ".conv2.groupnorm", prefix + ".conv2.conv", 3, 1, True, train=train,
                         x_needs_grad=r_needs_grad, y_stats=True)
        out = self.gn_conv(o, prefix + ".conv3.groupnorm", prefix + ".conv3.conv", 3, 1, True,
                           res=r, train=train, res_needs_grad=r_needs_grad, y_stats=True)
        return out

    def hr_module(self, xs, prefix, nb, train, outputs=None):
        """HighResolutionModule.forward (hr_util/hr3d.py:205-229)."""
        xs = self._branches(xs, prefix, nb, train)
        outs = []
        idx = list(range(nb) if outputs is None else outputs)
        # The fuse computations of the output branches read the same inputs and are independent of each other: in the
        # FORWARD pass output i >= 1 runs on branch stream i (forked here, joined below), so e.g. the element-wise
        # fuse_sum of output 0 overlaps the stride-2 convs of the others.  (Their backward closures stay on the main
        # stream: they accumulate into the shared input gradients in program order.)
        par = self.parallel_fuse and self.parallel_branches and len(idx) > 1
        self._s2d_share = {}
        for j in range(nb - 1):
            sib = ["%s.fuse_layers.%d.%d.0" % (prefix, i, j) for i in idx if i > j]
            if ops.USE_S2D_SHARE and len(sib) >= 2 and all(ops.s2d_eligible(xs[j], self.p[q + ".1.weight"]) for q in sib):
                # the stride-2 fuse convs out of branch j share one view of xhat (issued here, before the streams fork)
                self._s2d_share[id(xs[j])] = self._s2d_share_prepare(xs[j], len(sib))
        if par:
            dev = xs[0].buf.device
            main = torch.cuda.current_stream(dev)
            for x in xs:  # statistics every fuse conv may ask for, computed once, before the fork
                if self._stats_get(x) is None:
                    self._stats_put(x, ops.gn_stats(x, 8 if x.C >= 8 else 1))
            used = {i: self._fuse_stream(i, dev) for i in idx if i > 0}
            for st in used.values():  # fork everything first: a later wait_stream(main) would also wait for output 0
                st.wait_stream(main)
            if train and self.parallel_fuse_bwd:  # runs LAST among the fuse closures in backward
                def join_bwd():
                    cur = torch.cuda.current_stream(dev)
                    for st in used.values():
                        cur.wait_stream(st)
                self.tape.append(join_bwd)
        self._fuse_closures = []
        # Issue order: output 0, then the side outputs with the LONGEST stride-2 chain first (their first convs all want the
        # whole GPU and run one after the other: the chain with the most stages behind it should not be the last to start).
        # The backward closures and the fuse_sum closures keep the order of `idx` (gradient accumulation order unchanged).
        order = [i for i in idx if i == 0] + sorted((i for i in idx if i > 0), reverse=True) if par and FUSE_LONGEST_FIRST else idx
        ys, segs, fcl = {}, {}, {}
        base = len(self.tape)
        for i in order:
            st = used.get(i) if par else None
            t0, f0 = len(self.tape), len(self._fuse_closures)
            with (torch.cuda.stream(st) if st is not None else contextlib.nullcontext()):
                ys[i] = self._fuse_output(xs, prefix, nb, i, train)
            if st is not None and train and self.parallel_fuse_bwd:
                # backward of output i's fuse layers on the same stream; accumulation into the shared input gradients
                # is chained in program order by _grad_of / _wrote
                for k in range(t0, len(self.tape)):
                    self.tape[k] = self._on_stream(self.tape[k], st)
            segs[i], fcl[i] = self.tape[t0:], self._fuse_closures[f0:]
            del self.tape[t0:]
            del self._fuse_closures[f0:]
        assert len(self.tape) == base
        for i in idx:
            self.tape.extend(segs[i])
            self._fuse_closures.extend(fcl[i])
            outs.append(ys[i])
        self.tape.extend(self._fuse_closures)
        self._fuse_closures = []
        if par:
            for st in used.values():
                main.wait_stream(st)
            if train and self.parallel_fuse_bwd:  # runs FIRST: the side streams wait for the gradients of the outputs
                def fork_bwd():
                    cur = torch.cuda.current_stream(dev)
                    for st in used.values():
                        st.wait_stream(cur)
                self.tape.append(fork_bwd)
        return outs

    def _fuse_output(self, xs, prefix, nb, i, train):
        if True:
            same, low = [], []
            for j in range(nb):
                if j == i:
                    same.append(xs[j])
                elif j > i:
                    q = "%s.fuse_layers.%d.%d" % (prefix, i, j)
                    low.append(self.gn_conv(xs[j], q + ".0", q + ".1", 1, 1, False, train=train))
                else:
                    t = xs[j]
                    for k in range(i - j):
                        q = "%s.fuse_layers.%d.%d.%d" % (prefix, i, j, k)
                        grp = self._s2d_share.get(id(t)) if k == 0 else None
                        if grp is not None and grp["x"] is t:
                            t = self._gn_conv_s2d_shared(grp, q + ".0", q + ".1", k < i - j - 1, train)
                        else:
                            t = self.gn_conv(t, q + ".0", q + ".1", 3, 2, k < i - j - 1, train=train)
                    same.append(t)
            y = ops.fuse_sum(self.new(xs[i]), same, low, relu=True)
            y.relu_out = True
            if train:
                # the sum's own backward closures are appended by hr_module AFTER every output's conv closures, so in the
                # backward pass they all run first: each pass-through gradient (y_i -> xs[i]) is then still pending when the
                # first GroupNorm backward into xs[i] comes, and rides in it (_defer_add)
                self._fuse_closures.append(self._fuse_bwd(y, same, low))
            return y

    def _fuse_stream(self, i, device):
        """Stream of fuse output i >= 1.  Its priority is HIGHER than the main (capture) stream's: the stride-2 chains of the side
        outputs are sequences of small latency-bound kernels that end the module's critical path, while output 0's fuse_sum on
        the main stream is one long kernel of ~10^4 blocks — at equal or lower priority the chains' kernels queued behind all
        of its blocks (0.27 ms of a nearly empty GPU per module, profiles/r02_timeline_step.txt)."""
        if FUSE_PRIORITY == 0:
            return self._branch_stream(i, device)
        st = self._bstreams.get(("fuse", i, str(device)))
        if st is None:
            st = ops.named_stream(device, "fuse%d/%d" % (ops.LANE, i), priority=FUSE_PRIORITY)
            self._bstreams[("fuse", i, str(device))] = st
        return st

    def _branch_stream(self, b, device):
        st = self._bstreams.get((b, str(device)))
        if st is None:
            # shared by all engines of the process (see ops.named_stream)
            st = ops.named_stream(device, "branch%d/%d" % (ops.LANE, b), priority=BRANCH_PRIORITY)
            self._bstreams[(b, str(device))] = st
        return st

    def _branches(self, xs, prefix, nb, train):
        """One res_block per branch.  Branch 0 stays on the current stream; with parallel_branches the others fork onto
        side streams and join before the fuse layers.  The backward closures a branch records are replayed on the same
        side stream, bracketed by the mirrored fork / join."""
        if not self.parallel_branches or nb < 2:
            return [self.res_block(xs[b], "%s.branches.%d.0" % (prefix, b), train) for b in range(nb)]
        dev = xs[0].buf.device
        main = torch.cuda.current_stream(dev)
        side = [self._branch_stream(b, dev) for b in range(1, nb)]
        if train:  # runs LAST among this module's branch closures in backward: the current stream joins the sides
            def join_bwd():
                cur = torch.cuda.current_stream(dev)
                for st in side:
                    cur.wait_stream(st)
            self.tape.append(join_bwd)
        outs = [None] * nb
        for st in side:  # fork BEFORE branch 0 is issued: a later wait_stream(main) would also wait for branch 0
            st.wait_stream(main)
        for b in range(nb):
            if b == 0:
                outs[0] = self.res_block(xs[0], "%s.branches.0.0" % prefix, train)
                continue
            st = side[b - 1]
            t0 = len(self.tape)
            with torch.cuda.stream(st):
                outs[b] = self.res_block(xs[b], "%s.branches.%d.0" % (prefix, b), train)
            for i in range(t0, len(self.tape)):
                self.tape[i] = self._on_stream(self.tape[i], st)
        for st in side:
            main.wait_stream(st)
        if train:  # runs FIRST in backward: the sides wait for the fuse-layer gradients produced on the current stream
            def fork_bwd():
                cur = torch.cuda.current_stream(dev)
                for st in side:
                    st.wait_stream(cur)
            self.tape.append(fork_bwd)
        return outs

    @staticmethod
    def _on_stream(fn, st):
        def run():
            with torch.cuda.stream(st):
                fn()
        return run

    def _fuse_bwd(self, y, same, low):
        def bwd():
            g = self._g(y)
            if g is None:
                return
            for t in same:
                if t.grad is None and t.pending is None and not t.relu_out:
                    t.grad = g  # single consumer, no mask: alias instead of copying
                else:
                    self._defer_add(t, g)
            for t in low:
                gt, acc = self._grad_of(t)
                ops.upsample_bwd(g, gt, accumulate=acc)
                self._wrote(t)
        return bwd

    # ------------------------------------------------------------------ network
    def backbone(self, x, train):
        """HighResolution3DNet.forward (hr_util/hr3d.py:373-399) + HRNet3D.forward (backbones/hrnet3d.py:29-42)."""
        bb = self.pb + "backbone"
        x = self.res_block(x, bb + ".layer1", train, x_needs_grad=False)
        ys = [x]
        for s in (2, 3, 4):
            q = "%s.transition%d.%d.0" % (bb, s - 1, s - 1)
            new = self.gn_conv(ys[-1], q + ".0", q + ".1", 3, 2, True, train=train)
            outputs = [0] if (s == 4 and self.fuse == "top") else None  # y1..y3 of stage4 are never read ('top')
            ys = self.hr_module(ys + [new], "%s.stage%d.0" % (bb, s), s, train, outputs)
        if self.fuse == "top":
            if (self.pb + "final_conv.weight") in self.p:
                raise NotImplementedError("final_fuse='top' with a non-identity final_conv")
            return ys[0]
        # 'conat_conv': final_conv(cat(x0, up(x1), up(x2), up(x3))) == W0 x0 + sum_j up(Wj xj) + b  (1x1 conv and
        # trilinear interpolation commute); avoids materialising the concat.
        w, b = self.p[self.pb + "final_conv.weight"], self.p[self.pb + "final_conv.bias"]
        terms, c0 = [], 0
        f = ops.conat_forward(self.packs, ys, w, b, self.new(ys[0], C=w.shape[0]))  # one GEMM, concat assembled in shared memory
        for y in ys:
            t = None
            if f is None:
                t = self.new(y, C=w.shape[0])
                ops.conv_forward(self.packs, y, w, 1, t, ci0=c0, ci_n=y.C)
            terms.append((y, t, c0))
            c0 += y.C
        if f is None:
            f = ops.fuse_sum(self.new(ys[0], C=w.shape[0]), [terms[0][1]], [t for _, t, _ in terms[1:]], bias=b)
        if train:
            def bwd():
                g = self._g(f)
                if g is None:
                    return
                gb, accb = self._pgrad(self.pb + "final_conv.bias")
                gw, accw = self._pgrad(self.pb + "final_conv.weight")
                if not accw:
                    pass  # every input-channel slice is written below exactly once
                for i, (y, t, ci0) in enumerate(terms):
                    if i == 0:
                        gt = g
                    else:
                        gt = self.new(y, C=w.shape[0])
                        ops.upsample_bwd(g, gt)
                    # the bias gradient (sum of g over positions) rides in the full-resolution term's weight gradient as a
                    # ones channel of its GEMM N: g is not read once more for it (a 0.17 ms channel_sum pass before)
                    ops.conv_wgrad_async(y, gt, 1, 1, gw, accumulate=accw, ci0=ci0, bias_grad=(gb, accb) if i == 0 else None)
                    gy, acc = self._grad_of(y)
                    ops.conv_dgrad(self.packs, gt, w, 1, gy, mask=y if y.relu_out else None, accumulate=acc, ci0=ci0,
                                   ci_n=y.C)
                    self._wrote(y)
            self.tape.append(bwd)
        return f

    def head(self, f, train, reg_targets=None):
        """CenterHead.forward + SepHead.forward (pose_heads/center_head.py:232-238, :66-109); reg.0 and hm.0 read the
        same input and are merged into one N=64 GEMM."""
        p, ph = self.p, self.ph
        if ph + "shared_conv.1.weight" in p:
            f = self.gn_conv(f, ph + "shared_conv.0", ph + "shared_conv.1", 3, 1, True, train=train)
        q = ph + "tasks.0."
        w0 = _adjacent_cat(p[q + "reg.0.weight"], p[q + "hm.0.weight"])
        b0 = _adjacent_cat(p[q + "reg.0.bias"], p[q + "hm.0.bias"])
        hc = p[q + "reg.0.weight"].shape[0]
        wkey = ("merged_head", p[q + "reg.0.weight"].data_ptr(), p[q + "hm.0.weight"].data_ptr())
        wver = (p[q + "reg.0.weight"]._version, p[q + "hm.0.weight"]._version)
        t = self.new(f, C=2 * hc)
        t_reg, t_hm = t.channels(0, hc), t.channels(hc, hc)
        reg = self.new(f, C=self.R)
        hm = self.new(f, C=self.ncls)
        self._reg_out, self._reg2_weight = (reg, p[q + "reg.2.weight"]) if train else (None, None)  # see loss()
        # Training with known targets: the loss reads `reg` at the target voxels only, so the regression branch (reg.0 half of
        # the merged conv, reg.2) is evaluated on the units around them: within 1 voxel for the hidden layer (the 3x3x3 input
        # neighbourhood of a target), the target's own tile for reg.2.  The backward pass then MUST take the sparse paths.
        sparse_fwd = (train and reg_targets is not None and ops.USE_SPARSE_FWD and ops.USE_SPARSE_UNITS and hc == 32
                      and f.C % 32 == 0 and f.C > 80 and ops.reg_sparse_supported(p[q + "reg.2.weight"], reg_targets)
                      and ops.k3s1_eligible(f, f.C, 32) and ops.k3s1_eligible(t_reg, 32, 32)
                      and ops.k3s1_eligible(t_reg, 32, (self.R + 15) // 16 * 16))
        fwd_units = {}
        tg_pre = tg_ready = None
        if sparse_fwd and ops.USE_PREZERO:
            # the hidden layer's gradient buffer is taken now, and its regression half (written around the targets only by the
            # backward pass) is zero-filled on a side stream beside the head convolutions below
            tg_pre = self.new(t)
            tg_ready = ops.prezero(tg_pre.channels(0, hc))
        if sparse_fwd:
            u0 = ops.active_units(reg_targets, f, 0, "f0")
            u1 = fwd_units[1] = ops.active_units(reg_targets, f, 1, "f1")
            ops.conv_forward(self.packs, f, p[q + "hm.0.weight"], 1, t_hm, bias=p[q + "hm.0.bias"], relu=True)
            ops.conv_forward(self.packs, f, p[q + "reg.0.weight"], 1, t_reg, bias=p[q + "reg.0.bias"], relu=True, units=u1)
            ops.conv_forward(self.packs, t_reg, p[q + "reg.2.weight"], 1, reg, bias=p[q + "reg.2.bias"], units=u0)
        else:
            ops.conv_forward(self.packs, f, w0, 1, t, bias=b0, relu=True, key=wkey, version=wver)
            ops.conv_forward(self.packs, t_reg, p[q + "reg.2.weight"], 1, reg, bias=p[q + "reg.2.bias"])
        t.relu_out = True
        t_reg.relu_out = t_hm.relu_out = True
        ops.conv_forward(self.packs, t_hm, p[q + "hm.2.weight"], 1, hm, bias=p[q + "hm.2.bias"])
        if train:
            def bwd():
                if hm.grad is None or reg.grad is None:
                    return
                tg = tg_pre if tg_pre is not None else self.new(t)
                sp = self._reg_sparse
                self._reg_sparse = None
                reg_sparse = False
                if sparse_fwd and not (sp is not None and sp[0] is reg.grad and sp[1] is reg_targets):
                    raise lib.RtpError("forward(reg_targets=...) evaluated the regression branch around those targets only: "
                                       "loss() must be called with the same `ind` tensor before backward()")
                for name, tv, o, c0 in (("reg", t_reg, reg, 0), ("hm", t_hm, hm, hc)):
                    if (name == "reg" and sp is not None and sp[0] is reg.grad and ops.reg_sparse_supported(p[q + "reg.2.weight"], sp[1])):
                        reg_sparse = True
                        # the loss touches `reg` at the target voxels only: its gradient is zero elsewhere, and the whole
                        # backward of reg.2 (dgrad, weight and bias gradients) is a few hundred voxel neighbourhoods
                        gw, acc = self._pgrad(q + "reg.2.weight")
                        gb, accb = self._pgrad(q + "reg.2.bias")
                        if tg_ready is not None:
                            torch.cuda.current_stream(f.buf.device).wait_event(tg_ready)
                        ops.reg_head_bwd_sparse(reg.grad, tv, sp[1], p[q + "reg.2.weight"], tg.channels(c0, hc), gw, acc, gb, accb,
                                                prezeroed=tg_pre is not None)
                        continue
                    gw, acc = self._pgrad(q + name + ".2.weight")
                    ops.conv_wgrad_async(tv, o.grad, 3, 1, gw, accumulate=acc)
                    gb, accb = self._pgrad(q + name + ".2.bias")
                    ops.on_aux_stream(o.grad, lambda o=o, gb=gb, accb=accb: ops.channel_sum(o.grad, gb, accumulate=accb))
                    dy = o.grad
                    if dy.C8 == 1 and dy.n_stride >= 2 * dy.c_stride:  # gradient with a zeroed spare chunk (loss())
                        dy = P8(dy.N, 16, dy.Z, dy.Y, dy.X, buf=dy.buf, offset=dy.offset, n_stride=dy.n_stride,
                                c_stride=dy.c_stride)
                    ops.conv_dgrad(self.packs, dy, p[q + name + ".2.weight"], 1, tg.channels(c0, hc), mask=tv)
                if sp is not None and sp[0] is reg.grad and sp[2] and not reg_sparse:
                    raise lib.RtpError("loss() left the regression gradient undefined away from the targets, but the dense backward ran")
                gw_r, acc_r = self._pgrad(q + "reg.0.weight")
                gw_h, acc_h = self._pgrad(q + "hm.0.weight")
                # With the sparse regression gradient, the regression half of tg is non-zero only within one voxel of a target:
                # the two halves of the merged conv go out separately, and the regression half's weight gradient / dgrad visit
                # only the (sample, tile) units around the targets (within 1 / 2 voxels in the plane).
                split = (reg_sparse and ops.USE_SPARSE_UNITS and hc == 32 and f.C % 32 == 0 and f.C > 80
                         and ops.k3s1_eligible(tg.channels(0, hc), 32, 32))
                assert split or not sparse_fwd
                if split:
                    # (the forward pass listed the radius-1 units of the same targets already: sp[1] is reg_targets, checked above)
                    u1 = fwd_units[1] if sparse_fwd else ops.active_units(sp[1], f, 1, "r1")
                    u2 = ops.active_units(sp[1], f, 2, "r2")
                    ops.conv_wgrad_async(f, tg.channels(hc, hc), 3, 1, gw_h, accumulate=acc_h)
                    ops.conv_wgrad_async(f, tg.channels(0, hc), 3, 1, gw_r, accumulate=acc_r, units=u1)
                else:
                    ops.conv_wgrad_async(f, tg, 3, 1, gw_r, accumulate=acc_r, n0=0, more=((gw_h, acc_h, 0, hc),))
                for name, c0 in (("reg", 0), ("hm", hc)):
                    gb, accb = self._pgrad(q + name + ".0.bias")
                    ops.on_aux_stream(tg, lambda c0=c0, gb=gb, accb=accb: ops.channel_sum(tg.channels(c0, hc), gb, accumulate=accb))
                gf, accf = self._grad_of(f)
                fmask = f if f.relu_out else None
                if split:
                    ops.conv_dgrad(self.packs, tg.channels(hc, hc), p[q + "hm.0.weight"], 1, gf, mask=fmask, accumulate=accf)
                    ops.conv_dgrad(self.packs, tg.channels(0, hc), p[q + "reg.0.weight"], 1, gf, mask=fmask, accumulate=True, units=u2)
                else:
                    ops.conv_dgrad(self.packs, tg, w0, 1, gf, mask=fmask, accumulate=accf, key=wkey, version=wver)
                self._wrote(f)
            self.tape.append(bwd)
        return hm, reg

    # ------------------------------------------------------------------ entry points
    def begin(self):
        self.generation += 1
        ops.join_wgrad()  # side-stream work of a pass that was never back-propagated (e.g. the head's gradient pre-fill)
        self.pool.release_all()
        self.tape = []
        self.stats_cache = {}
        self._touched = set()

    def forward(self, x, train, reg_targets=None):
        """x: P8 input cube [N, in_ch, Z, Y, X].  Returns (hm, reg) raw head outputs as P8 tensors.
        reg_targets (training only): the int64 [N][M] target voxel indices the loss will gather the regression output at
        (CenterHead.loss, center_head.py:244-270).  When given, the regression branch of the head is evaluated only on the
        (sample, tile) units that contain those voxels — `reg` is then DEFINED ONLY THERE — and loss() / backward() must be
        called with the same indices (they take the matching sparse backward path)."""
        if x.buf.device.index != torch.cuda.current_device():
            # every launch goes to the CURRENT device's current stream (p8._stream): one process per GPU, or select the
            # device with torch.cuda.device(...) around the call
            raise lib.RtpError("input lives on %s but the current CUDA device is %d" % (x.buf.device, torch.cuda.current_device()))
        self.begin()
        f = self.backbone(x, train)
        return self.head(f, train, reg_targets)

    def loss(self, hm, reg, tgt_hm, ind, mask, cat, anno, with_grad=True, grad_scale=1.0):
        """CenterHead.loss (center_head.py:244-270).  Returns a device fp32 tensor
        [loss, hm_loss, loc_loss, num_pos, loc_loss_elem...]; seeds hm.grad / reg.grad when with_grad."""
        dev = hm.buf.device
        if self._cw is None or self._cw.device != dev:
            self._cw = torch.tensor(self.code_weights, dtype=torch.float32, device=dev)
        out = torch.empty(4 + self.R, dtype=torch.float32, device=dev)
        ws = ops.workspace(lib.load().rtp_head_loss_workspace_bytes(hm.N, self.ncls, hm.Z, hm.Y, hm.X), dev, "loss")
        flags = 0
        if with_grad:
            if hm.C8 == 1:
                # one spare, zeroed 8-channel chunk behind the heat-map gradient: the dgrad of hm.2 can then run on the
                # plane-streaming kernel (K = 16) instead of the generic one (see head.bwd); the loss kernel zero-fills it
                wide = self.new(hm, C=16)
                hm.grad = wide.channels(0, hm.C)
                dh = wide.struct()
            else:
                hm.grad = self.new(hm)
                dh = hm.grad.struct()
            reg.grad = self.new(reg)
            # The regression gradient is non-zero at the target voxels only.  When head.bwd is going to read it there only
            # (rtp_reg_head_bwd_sparse), the loss kernel does not zero-fill the rest of it (third entry: head.bwd checks it).
            w2 = self._reg2_weight
            sparse_only = bool(ops.USE_SPARSE_DREG and w2 is not None and reg is self._reg_out and ops.reg_sparse_supported(w2, ind))
            self._reg_sparse = (reg.grad, ind, sparse_only)
            if sparse_only:
                flags |= lib.RTP_LOSS_SPARSE_DREG
            dr = reg.grad.struct()
        else:
            dh = dr = lib.NULL_P8
        M = ind.shape[1]
        lib.call("rtp_head_loss_flags", hm.struct(), reg.struct(), self.ncls, self.R, tgt_hm.data_ptr(), ind.data_ptr(),
                 mask.data_ptr(), cat.data_ptr(), anno.data_ptr(), M, self.loss_weight, self._cw.data_ptr(),
                 float(grad_scale), out.data_ptr(), dh, dr, flags, ws.data_ptr(), _stream())
        return out

    def backward(self, grads):
        """Runs the tape in reverse; `grads`: dict name -> fp32 tensor receiving d loss / d param."""
        self.grads = grads
        self._touched = set()
        self._last_touch = {}
        self._ordered_grads = bool(self.parallel_fuse_bwd and self.parallel_fuse and self.parallel_branches)
        ms, fired = self._milestones, set()
        for i, fn in enumerate(reversed(self.tape)):
            self._closure_idx = i
            fn()
            if ms is not None:
                for k in ms.get(i, ()):
                    fired.add(k)
                    self._on_ready(k)
        ntape = len(self.tape)
        self.tape = []
        ops.join_wgrad()
        if self._grad_groups is not None:
            for k in range(len(self._grad_groups)):  # groups not reported on the way (first pass, or nothing touched)
                if k not in fired:
                    self._on_ready(k)
            learnt = {}
            for k, names in enumerate(self._grad_groups):
                idx = [self._last_touch[n] for n in names if n in self._last_touch]
                if idx:
                    learnt.setdefault(max(idx), []).append(k)
            self._milestones = learnt if ntape else None
        return self._touched

    def decode(self, hm, reg, voxel_xyz, range_xyz):
        """CenterHead.predict + post_processing (center_head.py:272-360) -> (index int32 [N,ncls], score, xyz)."""
        import ctypes as C
        dev = hm.buf.device
        idx = torch.empty((hm.N, self.ncls), dtype=torch.int32, device=dev)
        score = torch.empty((hm.N, self.ncls), dtype=torch.float32, device=dev)
        xyz = torch.empty((hm.N, self.ncls, self.R), dtype=torch.float32, device=dev)
        v = (C.c_float * 3)(*[float(a) for a in voxel_xyz])
        r = (C.c_float * 3)(*[float(a) for a in range_xyz])
        lib.call("rtp_decode", hm.struct(), reg.struct(), self.ncls, self.R, v, r, idx.data_ptr(), score.data_ptr(),
                 xyz.data_ptr(), _stream())
        return idx, score, xyz
