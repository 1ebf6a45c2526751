"""Drop-in boundary: the reference's det3d registry/builder API for the radar-pose path, backed by the B200 engine.

Mirrors (names, constructor kwargs, return types, error behaviour):
  det3d/utils/registry.py:6-78            Registry, build_from_cfg
  det3d/models/registry.py, builder.py    READERS/BACKBONES/HEADS/DETECTORS, build_reader/backbone/head/detector
  det3d/models/readers/radar_encoder.py   RadarFeatureNet
  det3d/models/backbones/hrnet3d.py       HRNet3D
  det3d/models/pose_heads/center_head.py  CenterHead (forward / loss / predict)
  det3d/models/detectors/radar_pose_net.py RadarPoseNet (extract_feat / forward)

`state_dict()` keys and shapes equal the reference's, so its checkpoints load unchanged.  All compute goes through
librtpose_b200.so; there is no PyTorch-op fallback (a CPU tensor or a missing library raises).
`install_as_det3d()` registers this module tree under the `det3d.*` names the reference's tools import.
"""
import inspect
import logging
import sys
import types
from collections import defaultdict

import torch
from torch import nn

from . import lib, ops, spec
from .engine import ARCH, Engine
from .p8 import P8


# ------------------------------------------------------------------------------------------------ registry
class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = dict()

    def __repr__(self):
        return self.__class__.__name__ + "(name={}, items={})".format(self._name, list(self._module_dict.keys()))

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if not inspect.isclass(cls):
            raise TypeError("module must be a class, but got {}".format(type(cls)))
        if cls.__name__ in self._module_dict:
            raise KeyError("{} is already registered in {}".format(cls.__name__, self.name))
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    assert isinstance(cfg, dict) and "type" in cfg
    assert isinstance(default_args, dict) or default_args is None
    args = dict(cfg)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError("{} is not in the {} registry".format(obj_type, registry.name))
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError("type must be a str or valid type, but got {}".format(type(obj_type)))
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    return obj_cls(**args)


READERS, BACKBONES, NECKS, HEADS, LOSSES, DETECTORS = (Registry(n) for n in
                                                       ("reader", "backbone", "neck", "head", "loss", "detector"))


def build(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_reader(cfg):
    return build(cfg, READERS)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_neck(cfg):
    return build(cfg, NECKS)


def build_head(cfg):
    return build(cfg, HEADS)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


# ------------------------------------------------------------------------------------------------ parameter trees
class _Node(nn.Module):
    """Name-only container: reproduces the reference's nested module names (Sequential indices, ModuleLists)."""


def _install(root, entries):
    for name, shape, init in entries:
        parts = name.split(".")
        m = root
        for p in parts[:-1]:
            if p not in m._modules:
                m.add_module(p, _Node())
            m = m._modules[p]
        m.register_parameter(parts[-1], nn.Parameter(spec.init_tensor(shape, init)))


def _named(module, prefix=""):
    return {prefix + k: v for k, v in module.named_parameters()}


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def _model_input(x, what):
    """The cube a detector is fed: the reference's fp32 [B,C,Z,Y,X] CUDA tensor, or an already packed P8 (what
    rtpose_b200.loader.CubeLoader yields), which is used as is."""
    if isinstance(x, P8):
        return x
    return _cuda_input(x, what).float().contiguous()


def _as_p8(x):
    return x if isinstance(x, P8) else P8.from_ncdhw(x)


def _x_key(x):
    return ("p8", x.N, x.C, x.Z, x.Y, x.X, x.n_stride, x.c_stride) if isinstance(x, P8) else tuple(x.shape)


def _x_static(x):
    return x.like() if isinstance(x, P8) else torch.empty_like(x)


def _x_copy(dst, src):
    if isinstance(src, P8):
        if (src.n_stride, src.c_stride, src.buf.numel(), src.offset) != (dst.n_stride, dst.c_stride, dst.buf.numel(), dst.offset):
            raise lib.RtpError("graphed step: a packed input must own its buffer (not a channel view)")
        dst.buf.copy_(src.buf)
    else:
        dst.copy_(src)


def _cuda_input(x, what):
    if not torch.is_tensor(x) or not x.is_cuda:
        raise lib.RtpError("%s must be a CUDA tensor: the rtpose_b200 path has no CPU fallback" % what)
    return x


# ------------------------------------------------------------------------------------------------ autograd bridge
class _Bridge(torch.autograd.Function):
    """Exposes one engine job (forward program + tape) to torch.autograd as a single node."""

    @staticmethod
    def forward(ctx, job, *tensors):
        ctx.job = job
        outs = job.run_forward()
        return outs if isinstance(outs, tuple) else (outs,)

    @staticmethod
    def backward(ctx, *gouts):
        grads = ctx.job.run_backward(gouts)
        return (None,) + tuple(grads)


class _ParamJob:
    """Shared plumbing: flat fp32 gradient buffer with one view per parameter."""

    def __init__(self, engine, params):
        self.engine, self.params = engine, params
        self.names = list(params)

    def grads_buffer(self):
        dev = next(iter(self.params.values())).device
        total = sum(p.numel() for p in self.params.values())
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        views, o = {}, 0
        for k, p in self.params.items():
            views[k] = flat[o:o + p.numel()].view(p.shape)
            o += p.numel()
        return flat, views

    def finish(self, flat, views, scale, touched):
        if scale is not None:
            flat.mul_(scale)
        return [views[k] if (k in touched and self.params[k].requires_grad) else None for k in self.names]


def _check_live(now, stamp):
    """An engine keeps ONE tape and one set of activation buffers: a second training forward (micro-batch accumulation,
    (loss_a + loss_b).backward()) recycles what the first one recorded.  Fail loudly instead of returning gradients of
    the wrong forward."""
    if now != stamp:
        raise lib.RtpError("backward() of a forward pass whose tape was overwritten by a later forward of the same module "
                           "(run forward -> backward pairs one at a time; accumulate gradients across them instead)")


class _StepJob(_ParamJob):
    """RadarPoseNet training step: input cube -> backbone -> head -> loss (+ gradient seeds) in one go.

    With `graph_state` (RadarPoseNet.cuda_graph = True) the whole step INCLUDING the backward pass is one captured CUDA
    graph over static input / target / gradient buffers: run_forward copies this iteration's tensors into them and
    replays; run_backward only hands out (a copy of) the gradients the replay already produced.  The graph is
    re-captured whenever a parameter, input or target tensor changes shape or a parameter moves in memory."""

    def __init__(self, engine, params, x, example, train, task_id=0, graph_state=None):
        super().__init__(engine, params)
        self.x, self.example, self.train, self.task_id = x, example, train, task_id
        self.graph_state = graph_state if train else None

    def _targets(self):
        ex, t, dev = self.example, self.task_id, self.x.device
        return (ex["hm"][t].to(dev, torch.float32).contiguous(), ex["ind"][t].to(dev, torch.int64).contiguous(),
                ex["mask"][t].to(dev, torch.uint8).contiguous(), ex["cat"][t].to(dev, torch.int64).contiguous(),
                ex["anno_pose"][t].to(dev, torch.float32).contiguous())

    def run_forward(self):
        if self.graph_state is not None:
            return self._run_graphed()
        e = self.engine
        tgt = self._targets()
        # training: the loss reads the regression map at tgt[1] (ind) only (Engine.forward, reg_targets)
        hm, reg = e.forward(_as_p8(self.x), self.train, reg_targets=tgt[1] if self.train else None)
        self.gen = e.generation
        return e.loss(hm, reg, *tgt, with_grad=self.train)

    def _run_graphed(self):
        from .graph import StepGraph
        st, e, tgt = self.graph_state, self.engine, self._targets()
        key = (tuple(p.data_ptr() for p in self.params.values()), _x_key(self.x), tuple(tuple(t.shape) for t in tgt))
        if st.get("key") != key:
            st.clear()
            st["x"], st["tgt"] = _x_static(self.x), tuple(torch.empty_like(t) for t in tgt)
            st["flat"], st["views"] = self.grads_buffer()

            def body():
                e.packs.refresh_async()  # the optimizer rewrites the weights between replays: repack inside the graph
                hm, reg = e.forward(_as_p8(st["x"]), True, reg_targets=st["tgt"][1])
                out = e.loss(hm, reg, *st["tgt"], with_grad=True)
                st["touched"] = e.backward(st["views"])
                return out
            _x_copy(st["x"], self.x)
            for d, t in zip(st["tgt"], tgt):
                d.copy_(t)
            st["graph"] = StepGraph(body, warmup=1).capture()
            st["key"] = key
        _x_copy(st["x"], self.x)
        for d, t in zip(st["tgt"], tgt):
            d.copy_(t)
        st["replays"] = st.get("replays", 0) + 1
        self.gen = st["replays"]
        return st["graph"]().clone()

    def run_backward(self, gouts):
        if self.graph_state is not None:
            st = self.graph_state
            _check_live(st.get("replays"), self.gen)
            flat = st["flat"].clone()  # the static buffer is overwritten by the next replay
            views, o = {}, 0
            for k, p in self.params.items():
                views[k] = flat[o:o + p.numel()].view(p.shape)
                o += p.numel()
            return self.finish(flat, views, gouts[0][0], st["touched"])
        _check_live(self.engine.generation, self.gen)
        flat, views = self.grads_buffer()
        touched = self.engine.backward(views)
        # d(out[0]) is the only differentiable element; its incoming gradient rescales everything linearly
        return self.finish(flat, views, gouts[0][0], touched)


class _BackboneJob(_ParamJob):
    def __init__(self, engine, params, x, train):
        super().__init__(engine, params)
        self.x, self.train = x, train

    def run_forward(self):
        e = self.engine
        e.begin()
        self.gen = e.generation
        self.f = e.backbone(P8.from_ncdhw(self.x), self.train)
        return self.f.to_ncdhw()

    def run_backward(self, gouts):
        e, f = self.engine, self.f
        _check_live(e.generation, self.gen)
        g = P8.from_ncdhw(gouts[0].contiguous())
        if f.relu_out:
            f.grad = e.new(f)
            ops.grad_add(g, f.grad, mask=f)
        else:
            f.grad = g
        flat, views = self.grads_buffer()
        touched = e.backward(views)
        return self.finish(flat, views, None, touched)


class _HeadJob(_ParamJob):
    """CenterHead.forward on an NCDHW feature tensor (input gradient returned as the first grad)."""

    def __init__(self, engine, params, x, train):
        super().__init__(engine, params)
        self.x, self.train = x, train

    def run_forward(self):
        e = self.engine
        e.begin()
        self.gen = e.generation
        self.fp = P8.from_ncdhw(self.x)
        self.hm, self.reg = e.head(self.fp, self.train)
        return self.hm.to_ncdhw(), self.reg.to_ncdhw()

    def run_backward(self, gouts):
        e = self.engine
        _check_live(e.generation, self.gen)
        self.hm.grad = P8.from_ncdhw(gouts[0].contiguous())
        self.reg.grad = P8.from_ncdhw(gouts[1].contiguous())
        flat, views = self.grads_buffer()
        touched = e.backward(views)
        gx = self.fp.grad.to_ncdhw() if self.fp.grad is not None else None
        return [gx] + self.finish(flat, views, None, touched)


class _LossJob:
    """CenterHead.loss on NCDHW head outputs: returns the loss vector; gradients w.r.t. (hm, reg)."""

    def __init__(self, engine, hm, reg, tgt, train):
        self.engine, self.hm_t, self.reg_t, self.tgt, self.train = engine, hm, reg, tgt, train

    def run_forward(self):
        e = self.engine
        self.hm, self.reg = P8.from_ncdhw(self.hm_t), P8.from_ncdhw(self.reg_t)
        return e.loss(self.hm, self.reg, *self.tgt, with_grad=self.train)

    def run_backward(self, gouts):
        s = gouts[0][0]
        return [self.hm.grad.to_ncdhw() * s, self.reg.grad.to_ncdhw() * s]


# ------------------------------------------------------------------------------------------------ modules
@READERS.register_module
class RadarFeatureNet(nn.Module):
    """det3d/models/readers/radar_encoder.py:7-17 — identity reader (the P8 packing happens inside the engine)."""

    def __init__(self, name="RadarFeatureNet"):
        super(RadarFeatureNet, self).__init__()
        self.name = name

    def forward(self, rdr_cube):
        return rdr_cube


@BACKBONES.register_module
class HRNet3D(nn.Module):
    """det3d/models/backbones/hrnet3d.py:8-56."""

    def __init__(self, backbone_cfg="hr_tiny_feat16_zyx_l4", feat_transform=None, **kwargs):
        super(HRNet3D, self).__init__()
        if backbone_cfg not in ARCH:
            raise KeyError(backbone_cfg)  # the reference raises KeyError from MODEL_CONFIGS[...] (hr3d.py:406-408)
        if feat_transform is not None:
            raise NotImplementedError("feat_transform is not part of the radar-pose configs (all pass None)")
        self.backbone_cfg = backbone_cfg
        self.final_fuse = kwargs["final_fuse"]
        self.final_conv_in, self.final_conv_out = kwargs["final_conv_in"], kwargs["final_conv_out"]
        self.with_feat_transform = False
        _install(self, spec.backbone_spec(backbone_cfg, self.final_conv_in, self.final_conv_out))
        if self.final_conv_in == self.final_conv_out:
            self.final_conv = nn.Identity()
        self.out_channels = ARCH[backbone_cfg][1][0] if self.final_fuse == "top" else self.final_conv_out
        self._engine = None

    def _eng(self):
        if self._engine is None:
            self._engine = Engine(self.backbone_cfg, self.final_fuse, _named(self), 3, 1, 0.0, [1.0] * 3,
                                  prefix_backbone="", prefix_head="unused.")
        return self._engine

    def forward(self, x_):
        x = _cuda_input(x_, "HRNet3D input").float()
        params = _named(self)
        job = _BackboneJob(self._eng(), params, x, torch.is_grad_enabled())
        self._eng().p = params
        return _Bridge.apply(job, *params.values())[0]


class DCNSepHead(nn.Module):
    """A 3-D-compatible definition of the reference's deformable head (center_head.py:111-163, `DCNSepHead`).

    The reference's class cannot be constructed (it passes bn= on to nn.Module.__init__, a TypeError: SURVEY F4) and, built
    from Conv2d / BatchNorm2d / a 2-D DeformConv, could not consume the 5-D feature map of the radar-pose backbone anyway.
    This module keeps its structure and parameter names and gives it a meaning on [B, C, Z, Y, X]:

      * every 2-D operator acts on the z-folded batch [B*Z, C, Y, X] — the (y, x) plane is the range-azimuth plane the
        deformable sampling is meant for; z slices are independent samples for the 2-D ops;
      * feature_adapt_cls / feature_adapt_reg: FeatureAdaption (1x1 offset predictor with zero-initialised weight, DCN v1
        3x3 with 4 deformable groups, ReLU), center_head.py:24-62;
      * cls_head: Conv2d 3x3 (C -> head_conv) + norm + ReLU + Conv2d 3x3 (head_conv -> num_cls, bias = init_bias).  The norm
        is GroupNorm(8) instead of BatchNorm2d: per-sample statistics, like every other norm of the path ('GCR'), so the
        result does not depend on how frames are sharded over ranks;
      * task_head: the ordinary SepHead on the un-folded reg feature (Conv3d 3x3x3, head_conv = 64), center_head.py:66-109.

    Enabled with CenterHead(dcn_head="fold_z"); dcn_head=True keeps raising TypeError like the reference.  Contractions run on
    the tcgen05 kernels (rtpose_b200.convfn, rtpose_b200.dcn with TENSOR_CORE)."""

    def __init__(self, in_channels, num_cls, heads, head_conv=64, final_kernel=3, init_bias=-2.19):
        super(DCNSepHead, self).__init__()
        from . import dcn
        if final_kernel != 3:
            raise NotImplementedError("DCNSepHead: final_kernel=%d" % final_kernel)
        self.feature_adapt_cls = dcn.FeatureAdaption(in_channels, in_channels, kernel_size=3, deformable_groups=4)
        self.feature_adapt_reg = dcn.FeatureAdaption(in_channels, in_channels, kernel_size=3, deformable_groups=4)
        self.heads, self.head_conv, self.num_cls = dict(heads), head_conv, num_cls
        # cls_head.{0: conv, 1: norm, 3: conv}: the reference's Sequential indices
        entries = [("cls_head.0.weight", (head_conv, in_channels, 3, 3), "conv"),
                   ("cls_head.0.bias", (head_conv,), "conv_bias:%d" % (in_channels * 9)),
                   ("cls_head.1.weight", (head_conv,), "ones"), ("cls_head.1.bias", (head_conv,), "zeros"),
                   ("cls_head.3.weight", (num_cls, head_conv, 3, 3), "conv"),
                   ("cls_head.3.bias", (num_cls,), "const:%r" % init_bias)]
        for name, (classes, num_conv) in self.heads.items():
            if num_conv != 2:
                raise NotImplementedError("SepHead with num_conv=%d" % num_conv)
            q = "task_head.%s" % name  # SepHead: every conv kaiming_init (normal, fan_out), bias 0 (center_head.py:96-99)
            entries += [(q + ".0.weight", (head_conv, in_channels, 3, 3, 3), "kaiming_fan_out"), (q + ".0.bias", (head_conv,), "zeros"),
                        (q + ".2.weight", (classes, head_conv, 3, 3, 3), "kaiming_fan_out"), (q + ".2.bias", (classes,), "zeros")]
        _install(self, entries)

    def _adapt_both(self, x5):
        """Both FeatureAdaption modules on the z-folded batch.  They read the same feature map, so the two 1x1 offset
        predictors run as ONE conv with 2 x 72 output channels (our kernel) and the z-fold of the input is done once; the
        DCN itself is rtpose_b200.dcn (tensor-core path)."""
        from . import convfn
        B, Cc, Z, Y, X = x5.shape
        fc, fr = self.feature_adapt_cls, self.feature_adapt_reg
        w = torch.cat([fc.conv_offset.weight, fr.conv_offset.weight], 0)
        b = torch.cat([fc.conv_offset.bias, fr.conv_offset.bias], 0)
        off5 = convfn.conv2d_as_3d(x5, w, b)                                                     # [B, 144, Z, Y, X]
        x2 = x5.permute(0, 2, 1, 3, 4).reshape(B * Z, Cc, Y, X)
        off2 = off5.permute(0, 2, 1, 3, 4).reshape(B * Z, off5.shape[1], Y, X)
        no = fc.conv_offset.weight.shape[0]
        outs = []
        for fa, o2 in ((fc, off2[:, :no]), (fr, off2[:, no:])):
            y2 = fa.relu(fa.conv_adaption(x2, o2))
            outs.append(y2.reshape(B, Z, -1, Y, X).permute(0, 2, 1, 3, 4))
        return outs

    def forward(self, x):
        from . import convfn
        x = _cuda_input(x, "DCNSepHead input").float()
        p = dict(self.named_parameters())
        center, regf = self._adapt_both(x)
        h = convfn.conv2d_as_3d(center, p["cls_head.0.weight"], p["cls_head.0.bias"])
        h = torch.relu(convfn.group_norm(h, 8, p["cls_head.1.weight"], p["cls_head.1.bias"]))
        ret = {}
        for name in self.heads:
            t = convfn.conv3d(regf, p["task_head.%s.0.weight" % name], p["task_head.%s.0.bias" % name], relu=True)
            ret[name] = convfn.conv3d(t, p["task_head.%s.2.weight" % name], p["task_head.%s.2.bias" % name])
        ret["hm"] = convfn.conv2d_as_3d(h, p["cls_head.3.weight"], p["cls_head.3.bias"])
        return ret


@HEADS.register_module
class CenterHead(nn.Module):
    """det3d/models/pose_heads/center_head.py:166-360."""

    def __init__(self, in_channels=128, tasks=[], dataset="cruw_pose", common_heads=dict(), logger=None, init_bias=-2.19,
                 share_conv_channel=64, num_hm_conv=2, weight=0.1, code_weights=[], dcn_head=False):
        super(CenterHead, self).__init__()
        if dcn_head and dcn_head != "fold_z":
            # the reference's DCNSepHead cannot be constructed either (TypeError at center_head.py:152)
            raise TypeError("dcn_head=True is not constructible in the reference (DCNSepHead passes bn= to "
                            "nn.Module.__init__); dcn_head='fold_z' selects the 3-D-compatible definition "
                            "(rtpose_b200.det3d_compat.DCNSepHead)")
        self.dcn_head = dcn_head or False
        num_classes = [len(_get(t, "class_names")) for t in tasks]
        if len(num_classes) != 1:
            raise NotImplementedError("the cruw_pose configs define exactly one task")
        self.class_names = [_get(t, "class_names") for t in tasks]
        self.weight, self.code_weights, self.dataset = weight, list(code_weights), dataset
        self.in_channels, self.num_classes = in_channels, num_classes
        self.logger = logger or logging.getLogger("CenterHead")
        heads = dict(common_heads)
        heads.update(dict(hm=(num_classes[0], num_hm_conv)))
        if list(heads) != ["reg", "hm"]:
            raise NotImplementedError("heads %s (the radar-pose path has common_heads={'reg': ...})" % list(heads))
        self.reg_channels = heads["reg"][0]
        if self.dcn_head:
            if in_channels != share_conv_channel:
                raise NotImplementedError("dcn_head with a shared_conv (in_channels != share_conv_channel)")
            self.shared_conv = nn.Identity()
            self.tasks = nn.ModuleList([DCNSepHead(share_conv_channel, num_classes[0], dict(common_heads), head_conv=64,
                                                   final_kernel=3, init_bias=init_bias)])
        else:
            _install(self, spec.head_spec(in_channels, share_conv_channel, heads, init_bias=init_bias))
            if in_channels == share_conv_channel:
                self.shared_conv = nn.Identity()
        self.sync_free_losses = False  # True: keep every returned loss term on the device (no .cpu() syncs)
        self._engine = None

    def _eng(self, params=None):
        if self._engine is None:
            self._engine = Engine("hr_tiny_feat32_zyx_l4_in32", "top", params or _named(self), self.reg_channels,
                                  self.num_classes[0], self.weight, self.code_weights, prefix_backbone="unused.",
                                  prefix_head="")
        return self._engine

    def forward(self, x, *kwargs):
        x = _cuda_input(x, "CenterHead input").float()
        if self.dcn_head:
            return [self.tasks[0](x)], x
        params = _named(self)
        e = self._eng(params)
        e.p = params
        hm, reg = _Bridge.apply(_HeadJob(e, params, x, torch.is_grad_enabled()), x, *params.values())
        return [{"reg": reg, "hm": hm}], x

    def _format_losses(self, out):
        R = self.reg_channels
        if self.sync_free_losses:
            hm_loss, elem = out[1].detach(), out[4:4 + R].detach()
        else:  # the reference hands back CPU copies (center_head.py:260)
            host = out.detach().cpu()
            hm_loss, elem = host[1], host[4:4 + R]
        ret = {"loss": out[0], "hm_loss": hm_loss, "loc_loss": out[2], "loc_loss_elem": elem, "num_positive": out[3].detach()}
        merged = defaultdict(list)
        for k, v in ret.items():
            merged[k].append(v)
        return merged

    def loss(self, example, preds_dicts, test_cfg, **kwargs):
        (preds,) = preds_dicts
        hm, reg = _cuda_input(preds["hm"], "preds['hm']"), _cuda_input(preds["reg"], "preds['reg']")
        dev = hm.device
        tgt = (example["hm"][0].to(dev, torch.float32).contiguous(), example["ind"][0].to(dev, torch.int64).contiguous(),
               example["mask"][0].to(dev, torch.uint8).contiguous(), example["cat"][0].to(dev, torch.int64).contiguous(),
               example["anno_pose"][0].to(dev, torch.float32).contiguous())
        e = self._eng()  # no e.begin() here: the head's forward tape (if any) must survive until backward
        (out,) = _Bridge.apply(_LossJob(e, hm.float().contiguous(), reg.float().contiguous(), tgt, torch.is_grad_enabled()), hm, reg)
        with torch.no_grad():  # the reference's loss() leaves the clamped sigmoid in preds_dict['hm'] (:248)
            preds["hm"] = torch.clamp(hm.detach().sigmoid(), min=1e-4, max=1 - 1e-4)
        return self._format_losses(out)

    def _keypoints(self, idx, score, xyz, test_cfg, metas):
        """post_processing (center_head.py:333-360) on the decoded batch (one D2H copy instead of one per sample)."""
        idx, score, xyz = idx.cpu(), score.cpu(), xyz.cpu()
        thr = _get(test_cfg, "score_threshold")
        ret = []
        for n in range(idx.shape[0]):
            kps = []
            if self.reg_channels == 3:
                for c in range(self.num_classes[0]):
                    s = float(score[n, c])
                    if s > thr:
                        kps.append((c, *[float(v) for v in xyz[n, c]], s))
            else:
                s = float(score[n, 0])
                pts = [float(v) for v in xyz[n, 0]]
                if s > thr:
                    kps.append((0, *pts[:3], s))
                for i in range(1, 15):
                    kps.append((i, *pts[3 * i:3 * (i + 1)], s))
            ret.append({"keypoints": kps, "metadata": metas[n] if metas is not None else None})
        return ret

    @torch.no_grad()
    def predict(self, example, preds_dicts, test_cfg, **kwargs):
        (preds,) = preds_dicts
        hm, reg = P8.from_ncdhw(_cuda_input(preds["hm"], "preds['hm']")), P8.from_ncdhw(preds["reg"])
        return self._predict_p8(hm, reg, test_cfg, example.get("meta") if isinstance(example, dict) else None)

    def _predict_p8(self, hm, reg, test_cfg, metas, engine=None):
        osf, vs, rng = _get(test_cfg, "out_size_factor"), _get(test_cfg, "voxel_size"), _get(test_cfg, "pc_range")
        voxel = (osf[2] * vs[0], osf[1] * vs[1], osf[0] * vs[2])
        idx, score, xyz = (engine or self._eng()).decode(hm, reg, voxel, rng[:3])
        return self._keypoints(idx, score, xyz, test_cfg, metas)


@DETECTORS.register_module
class RadarPoseNet(nn.Module):
    """det3d/models/detectors/radar_pose_net.py:9-46 (+ PoseNet.__init__, pose_net.py:12-33)."""

    def __init__(self, reader, backbone, neck, pose_head, sensor_type="rdr", train_cfg=None, test_cfg=None,
                 pretrained=None):
        super(RadarPoseNet, self).__init__()
        self.reader = build_reader(reader)
        self.backbone = build_backbone(backbone)
        if neck is not None:
            raise NotImplementedError("neck is None in every cruw_pose config")
        self.pose_head = build_head(pose_head)
        self.train_cfg, self.test_cfg, self.sensor_type = train_cfg, test_cfg, sensor_type
        self._engine = None
        # opt-in: run each training step (forward + loss + backward) as one replayed CUDA graph; see _StepJob
        self.cuda_graph = False
        self._graph_state = {}
        if pretrained is not None:
            try:
                ckpt = torch.load(pretrained, map_location="cpu")
                sd = ckpt.get("state_dict", ckpt)
                self.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}, strict=False)
                print("init weight from {}".format(pretrained))
            except Exception:  # the reference swallows this too (pose_net.py:37-41)
                print("no pretrained model at {}".format(pretrained))

    @property
    def with_neck(self):
        return False

    def _eng(self, params):
        if self._engine is None:
            b, h = self.backbone, self.pose_head
            self._engine = Engine(b.backbone_cfg, b.final_fuse, params, h.reg_channels, h.num_classes[0], h.weight,
                                  h.code_weights)
        self._engine.p = params
        return self._engine

    def extract_feat(self, data):
        return self.backbone(self.reader(data["rdr_tensor"]))

    def forward(self, example, return_loss=True, **kwargs):
        ex = {}
        ex.update(example[self.sensor_type])
        ex.update({"meta": example["meta"]})
        x = _model_input(self.reader(ex["rdr_tensor"]), "example['rdr']['rdr_tensor']")
        if getattr(self.pose_head, "dcn_head", False):
            # the deformable head is a module-level composition (DCNSepHead): backbone job -> head -> loss / predict,
            # as the reference's RadarPoseNet.forward does (radar_pose_net.py:36-46)
            if isinstance(x, P8):
                x = x.to_ncdhw()
            preds, _ = self.pose_head(self.backbone(x))
            if return_loss:
                return self.pose_head.loss(ex, preds, self.test_cfg)
            with torch.no_grad():
                return self.pose_head.predict(ex, preds, self.test_cfg)
        params = _named(self)
        e = self._eng(params)
        if return_loss:
            gs = self._graph_state if (self.cuda_graph and torch.is_grad_enabled()) else None
            (out,) = _Bridge.apply(_StepJob(e, params, x, ex, torch.is_grad_enabled(), graph_state=gs), *params.values())
            return self.pose_head._format_losses(out)
        with torch.no_grad():
            hm, reg = e.forward(_as_p8(x), False)
            return self.pose_head._predict_p8(hm, reg, self.test_cfg, ex["meta"], engine=e)


# ------------------------------------------------------------------------------------------------ det3d aliasing
def _real_or_shell(name, attach_to=None):
    """The real det3d module `name` when the reference is on sys.path and it imports; otherwise a shell module whose
    __path__ points at the real package directory when that exists (so its untouched sub-modules stay importable) or is
    empty (no reference on sys.path at all).  Never replaces a module that is already in sys.modules."""
    import importlib
    import importlib.util
    m = sys.modules.get(name)
    if m is not None:
        return m
    spec = None
    try:
        spec = importlib.util.find_spec(name)
    except Exception:
        spec = None
    if spec is not None:
        try:
            return importlib.import_module(name)
        except Exception:
            sys.modules.pop(name, None)  # e.g. det3d.models needs spconv / pycocotools: fall through to a shell
    m = types.ModuleType(name)
    m.__path__ = list(spec.submodule_search_locations or []) if spec is not None and spec.submodule_search_locations else []
    m.__package__ = name
    if spec is not None:
        m.__spec__ = spec
    sys.modules[name] = m
    if "." in name:
        parent = sys.modules.get(name.rsplit(".", 1)[0])
        if parent is not None:
            setattr(parent, name.rsplit(".", 1)[1], m)
    return m


def install_as_det3d():
    """Makes `from det3d.models import build_detector` (tools/train.py:22, tools/test.py:20) and the builder / registry
    names resolve to this implementation WITHOUT hiding the rest of the reference: when the real `det3d` package is
    importable (the maintainer's case) it is imported first and only the model-builder attributes are patched onto
    `det3d.models`, `det3d.models.builder`, `det3d.models.registry` and `det3d.builder`; `det3d.utils`, `det3d.torchie`,
    `det3d.datasets`, ... stay the reference's own modules.  Shell modules are created only for names that cannot be
    imported (no reference on sys.path, or `det3d.models` failing on its optional dependencies)."""
    me = sys.modules[__name__]
    _real_or_shell("det3d")
    exported = ("build_detector", "build_backbone", "build_head", "build_reader", "build_neck", "build", "Registry",
                "build_from_cfg", "READERS", "BACKBONES", "NECKS", "HEADS", "LOSSES", "DETECTORS", "RadarPoseNet",
                "HRNet3D", "CenterHead", "RadarFeatureNet")
    for n in ("det3d", "det3d.models", "det3d.models.builder", "det3d.models.registry", "det3d.builder"):
        m = _real_or_shell(n)
        for k in exported:
            setattr(m, k, getattr(me, k))
    u = _real_or_shell("det3d.utils")  # the reference's own module when present: only fill in what it lacks
    for k in ("Registry", "build_from_cfg"):
        if not hasattr(u, k):
            setattr(u, k, getattr(me, k))
    # `from det3d.ops.dcn import DeformConv` (pose_heads/center_head.py:18) and det3d/ops/dcn/__init__.py's other names:
    # the reference's extension is not built on sm_100, so these always resolve to the B200 kernels
    from . import dcn
    ops_mod = _real_or_shell("det3d.ops")
    ops_mod.dcn = dcn
    sys.modules["det3d.ops.dcn"], sys.modules["det3d.ops.dcn.deform_conv"] = dcn, dcn
    sys.modules["det3d"].ops = ops_mod
    return sys.modules["det3d"]
