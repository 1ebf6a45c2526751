"""Parameter inventory of the HRRadarPose path: names, shapes and initialisers, mirroring what the reference's
constructors register so that `state_dict()` keys/shapes are identical and reference
checkpoints load unchanged.

Reference construction order followed here: HighResolution3DNet.__init__ (hr_util/hr3d.py:234-284: layer1,
transition1, stage2, transition2, stage3, transition3, stage4), HRNet3D.__init__ (backbones/hrnet3d.py:11-27:
backbone, final_conv), CenterHead.__init__ (pose_heads/center_head.py:168-229: shared_conv, tasks[SepHead]).
"""
import math

import torch

from .engine import ARCH


def backbone_spec(arch, final_in, final_out):
    """[(name, shape, init)] for HRNet3D; init in {'conv', 'conv_bias:<fan_in>', 'ones', 'zeros'}."""
    inp, ch = ARCH[arch]
    spec = []

    def gn(p, n):
        spec.append((p + ".weight", (n,), "ones"))
        spec.append((p + ".bias", (n,), "zeros"))

    def conv(p, cout, cin, k, bias=False):
        spec.append((p + ".weight", (cout, cin, k, k, k), "conv"))
        if bias:
            spec.append((p + ".bias", (cout,), "conv_bias:%d" % (cin * k ** 3)))

    def block(p, cin, cout):  # ResNetBlock (hr_util/common.py:98-136)
        if cin != cout:
            conv(p + ".conv1", cout, cin, 1, bias=True)
        for name in ("conv2", "conv3"):
            gn(p + "." + name + ".groupnorm", cout)
            conv(p + "." + name + ".conv", cout, cout, 3)

    bb = "backbone"
    block(bb + ".layer1", inp, ch[0])
    for s in (2, 3, 4):
        p = "%s.transition%d.%d.0" % (bb, s - 1, s - 1)  # hr3d.py:286-331: new branch from the last one
        gn(p + ".0", ch[s - 2])
        conv(p + ".1", ch[s - 1], ch[s - 2], 3)
        sp = "%s.stage%d.0" % (bb, s)
        for b in range(s):
            block("%s.branches.%d.0" % (sp, b), ch[b], ch[b])
        for i in range(s):  # hr3d.py:135-200
            for j in range(s):
                if j > i:
                    q = "%s.fuse_layers.%d.%d" % (sp, i, j)
                    gn(q + ".0", ch[j])
                    conv(q + ".1", ch[i], ch[j], 1)
                elif j < i:
                    for k in range(i - j):
                        q = "%s.fuse_layers.%d.%d.%d" % (sp, i, j, k)
                        gn(q + ".0", ch[j])
                        conv(q + ".1", ch[i] if k == i - j - 1 else ch[j], ch[j], 3)
    if final_in != final_out:
        conv("final_conv", final_out, final_in, 1, bias=True)
    return spec


def head_spec(in_channels, share, heads, init_bias=-2.19, head_conv=32):
    """heads: ordered dict name -> (classes, num_conv) exactly as SepHead receives it (common_heads then 'hm')."""
    spec = []
    if in_channels != share:
        spec.append(("shared_conv.0.weight", (in_channels,), "ones"))
        spec.append(("shared_conv.0.bias", (in_channels,), "zeros"))
        spec.append(("shared_conv.1.weight", (share, in_channels, 3, 3, 3), "conv"))
    for name, (classes, num_conv) in heads.items():
        if num_conv != 2:
            raise NotImplementedError("SepHead with num_conv=%d (the cruw_pose configs use 2)" % num_conv)
        q = "tasks.0.%s" % name
        is_hm = "hm" in name
        # center_head.py:94-99: hm -> default conv init + last bias = init_bias; others -> kaiming_init on every conv
        wi = "conv" if is_hm else "kaiming_fan_out"
        spec.append((q + ".0.weight", (head_conv, share, 3, 3, 3), wi))
        spec.append((q + ".0.bias", (head_conv,), ("conv_bias:%d" % (share * 27)) if is_hm else "zeros"))
        spec.append((q + ".2.weight", (classes, head_conv, 3, 3, 3), wi))
        spec.append((q + ".2.bias", (classes,), ("const:%r" % init_bias) if is_hm else "zeros"))
    return spec


def init_tensor(shape, init):
    t = torch.empty(shape, dtype=torch.float32)
    if init == "ones":
        return t.fill_(1.0)
    if init == "zeros":
        return t.zero_()
    if init == "conv":  # nn.Conv3d default: kaiming_uniform_(a=sqrt(5))
        torch.nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        return t
    if init.startswith("conv_bias:"):
        bound = 1.0 / math.sqrt(int(init.split(":")[1]))
        return t.uniform_(-bound, bound)
    if init == "kaiming_fan_out":  # torchie/cnn/weight_init.py:32-45 (normal, fan_out, relu)
        torch.nn.init.kaiming_normal_(t, a=0, mode="fan_out", nonlinearity="relu")
        return t
    if init.startswith("const:"):
        return t.fill_(float(init.split(":")[1]))
    raise ValueError(init)
