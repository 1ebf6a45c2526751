"""TEST INFRASTRUCTURE — CPU restatement (the "oracle") of the RT-Pose HRRadarPose hot path.

Plain functional torch/numpy, fp32, keyed by the reference's `state_dict` names, so it can travel to the
GPU box (where /root/reference does not exist).  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it; nothing
under `rtpose_b200/` does.

Parity pin: the reference ships no tests or golden vectors, so this restatement is pinned
against outputs of the reference's own unmodified modules executed in the build container
(checked by `tests/test_oracle_golden.py`, `tests/test_evaluation.py`, `tests/test_abi_and_host.py`):
  model forward / loss / gradients / decode   oracle/make_golden.py        -> tests/golden/*_g*x*x*.npz
  ingest (ROI indices, crop, normalise, clamp)  oracle/make_ingest_golden.py -> tests/golden/ingest_golden.npz
  target assignment (both label layouts)        oracle/make_target_golden.py -> tests/golden/targets_golden.npz
  optimizer step (clip + decay + Adam)          oracle/make_optim_golden.py  -> tests/golden/optim_golden.npz
  one-cycle schedule                            oracle/make_sched_golden.py  -> tests/golden/one_cycle_golden.json
  evaluation (PJPE / MPJPE)                     oracle/make_eval_golden.py   -> tests/golden/eval_golden.json
The arithmetic
itself (conv3d / group_norm / trilinear interpolate) lives in PyTorch, the reference's pinned third-party
dependency (requirements-torch.txt:1, torch==2.0.1; here torch 2.11) — both sides call the same ATen ops.

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

# det3d/models/backbones/hrnet3D_config.py:85-177 — (input planes, per-branch channels) of the arch tables
ARCH = {
    "hr_tiny_feat32_zyx_l4": (1, [32, 32, 64, 64]),
    "hr_tiny_feat32_zyx_l4_in32": (32, [32, 32, 64, 64]),
    "hr_tiny_feat64_zyx_l4_in64": (64, [64, 64, 128, 128]),
}

# configs/cruw_pose/*.py `model=` blocks reduced to what the path needs
CONFIGS = {
    "hr3d": dict(arch="hr_tiny_feat32_zyx_l4", final_in=32, final_out=32, fuse="top", head_in=32, share=32,
                 reg=3, ncls=15, weight=0.2, code_weights=[1.0, 1.5, 2.0], norm=(150000.0, 200000.0), in_ch=1),
    "hr3d_one_hm": dict(arch="hr_tiny_feat32_zyx_l4", final_in=192, final_out=128, fuse="conat_conv", head_in=128,
                        share=128, reg=45, ncls=1, weight=0.5, code_weights=[1.0] * 45, norm=(150000.0, 200000.0),
                        in_ch=1),
    "hr3d_one_hm_doppler": dict(arch="hr_tiny_feat32_zyx_l4_in32", final_in=192, final_out=128, fuse="conat_conv",
                                head_in=128, share=128, reg=45, ncls=1, weight=0.5, code_weights=[1.0] * 45,
                                norm=(0.0, 10.0), in_ch=32),
    "hr3d_one_hm_doppler_phase": dict(arch="hr_tiny_feat64_zyx_l4_in64", final_in=384, final_out=256,
                                      fuse="conat_conv", head_in=256, share=256, reg=45, ncls=1, weight=0.7,
                                      code_weights=[1.0] * 45, norm=None, in_ch=64),
}

# configs/cruw_pose/hr3d.py:31-39,101,127-129 — ROI geometry
VOXEL_SIZE = (0.0453125, 0.15703125, 0.3625)  # x, y, z
PC_RANGE = (0.7703125, -5.0250000000000234, -1.0875000000000021)  # x, y, z minima of roi1
ROI_IDX = (13, 28, 32, 95, 17, 176)  # z, y, x inclusive index ranges (cruw_pose.py:125-146 evaluated on :38-40)


# ------------------------------------------------------------------------------------------------ ingest
def ingest_cube(raw, norm):
    """det3d/datasets/cruw_pose/cruw_pose.py:167-185 (`get_cube`) + pipelines/pose.py:163-172 channel packing.

    raw: float16 numpy [D,32,128,256] ('dzyx_real') or [32,128,256] ('zyx_real').  Returns fp32 [C,16,64,160].
    """
    a = raw.astype(np.float32)
    z0, z1, y0, y1, x0, x1 = ROI_IDX
    a = a[..., z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
    start, scale = float(norm[0]), float(norm[1]) - float(norm[0])
    a = (a - start) / scale
    a[a < 0.0] = 0.0
    if a.ndim < 4:
        a = a[None]
    return a


def ingest_cube_phase(raw):
    """cruw_pose.py:188-194 (`get_cube_phase`: crop only) + pose.py:168-170 ((2,D,Z,Y,X)->(2D,Z,Y,X))."""
    a = raw.astype(np.float32)
    z0, z1, y0, y1, x0, x1 = ROI_IDX
    a = a[:, :, z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
    return a.reshape(-1, *a.shape[2:])


# ------------------------------------------------------------------------------------------------ targets
def gaussian3d(diameter, sigma):
    """det3d/core/utils/center_utils.py:67-72 (note the (2 sigma^2)^(3/2) denominator)."""
    m = (diameter - 1.0) / 2.0
    z, y, x = np.ogrid[-m:m + 1, -m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y + z * z) / (2 * sigma * sigma) ** (3 / 2))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_gaussian3d(hm, center_xyz, radius):
    """center_utils.py:74-91 — elementwise-max splat of the gaussian, clipped at the volume border."""
    d = 2 * radius + 1
    g = gaussian3d(d, d / 6)
    x, y, z = int(center_xyz[0]), int(center_xyz[1]), int(center_xyz[2])
    Z, Y, X = hm.shape
    fx, rx = min(x, radius), min(X - x, radius + 1)
    fy, ry = min(y, radius), min(Y - y, radius + 1)
    fz, rz = min(z, radius), min(Z - z, radius + 1)
    mh = hm[z - fz:z + rz, y - fy:y + ry, x - fx:x + rx]
    mg = g[radius - fz:radius + rz, radius - fy:radius + ry, radius - fx:radius + rx]
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg, out=mh)
    return hm


def assign_targets(pose_xyz, grid_zyx, one_hm, radius=2, pc_range=PC_RANGE, voxel=VOXEL_SIZE):
    """One frame, one pose (max_poses=1).  pose_xyz: [15,3] metres.

    one_hm=True : pipelines/pose.py:385-452 (`AssignLabelPose2`): 1 heatmap, M=1, anno_pose[1,45].
    one_hm=False: pipelines/pose.py:186-255 (`AssignLabelPose`) : 15 heatmaps, M=15, anno_pose[15,3].
    """
    Z, Y, X = grid_zyx
    pose_xyz = np.asarray(pose_xyz, dtype=np.float64)
    # radar_range is float32 in the reference (pose.py:190,389); coordinates are python floats
    rr = np.array([pc_range[2], pc_range[1], pc_range[0]], dtype=np.float32)  # z,y,x minima
    if one_hm:
        hm = np.zeros((1, Z, Y, X), np.float32)
        anno = np.zeros((1, 45), np.float32)
        ind = np.zeros((1,), np.int64)
        mask = np.zeros((1,), np.uint8)
        cat = np.zeros((1,), np.int64)
        ct = []
        for i in range(15):
            x, y, z = pose_xyz[i]
            ct += [(x - rr[2]) / voxel[0], (y - rr[1]) / voxel[1], (z - rr[0]) / voxel[2]]
        ct = np.array(ct, dtype=np.float32)
        ci = ct.astype(np.int32)[:3]
        if 0 <= ci[0] < X and 0 <= ci[1] < Y and 0 <= ci[2] < Z:
            draw_gaussian3d(hm[0], ci, radius)
            ind[0] = ci[2] * Y * X + ci[1] * X + ci[0]
            mask[0] = 1
            anno[0] = (ct.reshape(-1, 3) - ci[None, :].astype(np.float32)).flatten()
    else:
        hm = np.zeros((15, Z, Y, X), np.float32)
        anno = np.zeros((15, 3), np.float32)
        ind = np.zeros((15,), np.int64)
        mask = np.zeros((15,), np.uint8)
        cat = np.zeros((15,), np.int64)
        for k in range(15):
            x, y, z = pose_xyz[k]
            ct = np.array([(x - rr[2]) / voxel[0], (y - rr[1]) / voxel[1], (z - rr[0]) / voxel[2]], dtype=np.float32)
            ci = ct.astype(np.int32)
            if not (0 <= ci[0] < X and 0 <= ci[1] < Y and 0 <= ci[2] < Z):
                continue
            draw_gaussian3d(hm[k], ci, max(radius, 1))
            cat[k] = k
            ind[k] = ci[2] * Y * X + ci[1] * X + ci[0]
            mask[k] = 1
            anno[k] = ct - ci.astype(np.float32)
    return dict(hm=hm, anno_pose=anno, ind=ind, mask=mask, cat=cat)


# ------------------------------------------------------------------------------------------------ backbone
def _gn(x, sd, key):
    c = x.shape[1]
    return F.group_norm(x, 8 if c >= 8 else 1, sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def res_block(x, sd, p):
    """hr_util/common.py:138-148 — r=conv1(x); o=relu(conv(gn(r))); o=conv(gn(o)); relu(o+r)."""
    if p + ".conv1.weight" in sd:
        r = F.conv3d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    else:
        r = x
    o = F.relu(F.conv3d(_gn(r, sd, p + ".conv2.groupnorm"), sd[p + ".conv2.conv.weight"], padding=1))
    o = F.conv3d(_gn(o, sd, p + ".conv3.groupnorm"), sd[p + ".conv3.conv.weight"], padding=1)
    return F.relu(o + r)


def _gn_conv(x, sd, p, stride, k3, relu):
    o = F.conv3d(_gn(x, sd, p + ".0"), sd[p + ".1.weight"], stride=stride, padding=1 if k3 else 0)
    return F.relu(o) if relu else o


def hr_module(xs, sd, p, nb):
    """hr_util/hr3d.py:205-229 (`HighResolutionModule.forward`), fuse layers built at :135-200."""
    xs = [res_block(xs[b], sd, "%s.branches.%d.0" % (p, b)) for b in range(nb)]
    outs = []
    for i in range(nb):
        y = None
        for j in range(nb):
            if j == i:
                t = xs[j]
            elif j > i:
                t = _gn_conv(xs[j], sd, "%s.fuse_layers.%d.%d" % (p, i, j), 1, False, False)
                t = F.interpolate(t, size=xs[i].shape[2:], mode="trilinear", align_corners=True)
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _gn_conv(t, sd, "%s.fuse_layers.%d.%d.%d" % (p, i, j, k), 2, True, k < i - j - 1)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hr_backbone(x, sd, p="backbone.backbone"):
    """hr_util/hr3d.py:373-399 (`HighResolution3DNet.forward`) with transitions :286-331."""
    x = res_block(x, sd, p + ".layer1")
    ys = [x]
    for s in (2, 3, 4):
        new = _gn_conv(ys[-1], sd, "%s.transition%d.%d.0" % (p, s - 1, s - 1), 2, True, True)
        ys = hr_module(ys + [new], sd, "%s.stage%d.0" % (p, s), s)
    return ys


def hrnet3d(x, sd, fuse):
    """det3d/models/backbones/hrnet3d.py:29-42."""
    ys = hr_backbone(x, sd)
    if fuse == "top":
        f = ys[0]
    else:
        size = ys[0].shape[2:]
        f = torch.cat([ys[0]] + [F.interpolate(y, size=size, mode="trilinear", align_corners=True) for y in ys[1:]], 1)
    if "backbone.final_conv.weight" in sd:
        f = F.conv3d(f, sd["backbone.final_conv.weight"], sd["backbone.final_conv.bias"])
    return f


# ------------------------------------------------------------------------------------------------ head
def center_head(f, sd, p="pose_head"):
    """det3d/models/pose_heads/center_head.py:232-238 + SepHead :66-109 (conv3+b, relu, conv3+b per head)."""
    if p + ".shared_conv.1.weight" in sd:
        f = F.relu(F.conv3d(_gn(f, sd, p + ".shared_conv.0"), sd[p + ".shared_conv.1.weight"], padding=1))
    out = {}
    for h in ("reg", "hm"):
        q = "%s.tasks.0.%s" % (p, h)
        t = F.relu(F.conv3d(f, sd[q + ".0.weight"], sd[q + ".0.bias"], padding=1))
        out[h] = F.conv3d(t, sd[q + ".2.weight"], sd[q + ".2.bias"], padding=1)
    return out


def _gather(feat, ind):
    """det3d/core/utils/center_utils.py:103-117 (`_transpose_and_gather_feat`)."""
    b, c = feat.shape[:2]
    f = feat.reshape(b, c, -1).permute(0, 2, 1)
    return f.gather(1, ind.unsqueeze(2).expand(-1, -1, c))


def head_loss(preds, tgt, weight, code_weights):
    """center_head.py:244-270 + losses/centernet_loss.py:17-24 (RegLoss) and :34-54 (FastFocalLoss).

    preds: {'hm','reg'} raw logits NCDHW; tgt: dict of batched tensors hm,ind,mask,cat,anno_pose.
    """
    p = torch.clamp(torch.sigmoid(preds["hm"]), min=1e-4, max=1 - 1e-4)
    mask = tgt["mask"].float()
    neg = (torch.log(1 - p) * p.pow(2) * (1 - tgt["hm"]).pow(4)).sum()
    pos_pix = _gather(p, tgt["ind"])
    pos_pred = pos_pix.gather(2, tgt["cat"].unsqueeze(2))
    num_pos = mask.sum()
    pos = (torch.log(pos_pred) * (1 - pos_pred).pow(2) * mask.unsqueeze(2)).sum()
    hm_loss = -neg if num_pos == 0 else -(pos + neg) / num_pos
    pred = _gather(preds["reg"], tgt["ind"])
    m = mask.unsqueeze(2)
    l1 = (pred * m - tgt["anno_pose"] * m).abs() / (m.sum() + 1e-4)
    reg_loss = l1.transpose(2, 0).sum(dim=2).sum(dim=1)
    loc_loss = (reg_loss * reg_loss.new_tensor(code_weights)).sum()
    return dict(loss=hm_loss + weight * loc_loss, hm_loss=hm_loss, loc_loss=loc_loss, loc_loss_elem=reg_loss,
                num_positive=mask.sum())


def decode(hm_logits, reg, score_threshold=0.0, voxel=VOXEL_SIZE, pc_range=PC_RANGE):
    """center_head.py:272-360 (`predict` + `post_processing`): sigmoid, per-class argmax over Z*Y*X (first
    index wins ties), metric coordinates (idx + reg) * voxel + range.  Returns per sample a list of
    (label, x, y, z, score) and the integer flat indices used.
    """
    b, ncls = hm_logits.shape[:2]
    Z, Y, X = hm_logits.shape[2:]
    hm = torch.sigmoid(hm_logits.float()).reshape(b, ncls, -1)
    rg = reg.float().reshape(b, reg.shape[1], -1)
    nk = reg.shape[1] // 3
    out, idxs = [], []
    for n in range(b):
        kps, ii = [], []
        for c in range(ncls):
            ind = int(torch.argmax(hm[n, c]))
            score = float(hm[n, c, ind])
            z, r = divmod(ind, Y * X)
            y, x = divmod(r, X)
            ii.append(ind)
            pts = []
            for i in range(nk):
                xs = (torch.tensor(float(x)) + rg[n, 3 * i, ind]) * 1 * voxel[0] + pc_range[0]
                ys = (torch.tensor(float(y)) + rg[n, 3 * i + 1, ind]) * 1 * voxel[1] + pc_range[1]
                zs = (torch.tensor(float(z)) + rg[n, 3 * i + 2, ind]) * 1 * voxel[2] + pc_range[2]
                pts += [float(xs), float(ys), float(zs)]
            if nk == 1:
                if score > score_threshold:
                    kps.append((c, pts[0], pts[1], pts[2], score))
            else:
                if score > score_threshold:
                    kps.append((0, pts[0], pts[1], pts[2], score))
                for i in range(1, 15):
                    kps.append((i, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], score))
        out.append(kps)
        idxs.append(ii)
    return out, idxs


# ------------------------------------------------------------------------------------------------ whole path
def forward(x, sd, cfg):
    """det3d/models/detectors/radar_pose_net.py:26-46: reader (identity) -> backbone -> head."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return center_head(hrnet3d(x, sd, c["fuse"]), sd)


def forward_loss(x, sd, cfg, tgt):
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return head_loss(forward(x, sd, c), tgt, c["weight"], c["code_weights"])


# ------------------------------------------------------------------------------------------------ weights
def state_dict_spec(cfg):
    """Names and shapes of the reference state_dict, in the reference's order."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    inp, ch = ARCH[c["arch"]]
    spec = []

    def gn(p, n):
        spec.append((p + ".weight", (n,)))
        spec.append((p + ".bias", (n,)))

    def block(p, cin, cout):
        if cin != cout:
            spec.append((p + ".conv1.weight", (cout, cin, 1, 1, 1)))
            spec.append((p + ".conv1.bias", (cout,)))
        for k in ("conv2", "conv3"):
            gn(p + "." + k + ".groupnorm", cout)
            spec.append((p + "." + k + ".conv.weight", (cout, cout, 3, 3, 3)))

    bb = "backbone.backbone"
    block(bb + ".layer1", inp, ch[0])
    trans, stages = [], []
    for s in (2, 3, 4):
        p = "%s.transition%d.%d.0" % (bb, s - 1, s - 1)
        cin = ch[s - 2]
        trans.append((p, cin, ch[s - 1]))
    for s in (2, 3, 4):
        stages.append(s)
    # reference registration order: layer1, transition1, stage2, transition2, stage3, transition3, stage4
    for s in (2, 3, 4):
        p, cin, cout = trans[s - 2]
        gn(p + ".0", cin)
        spec.append((p + ".1.weight", (cout, cin, 3, 3, 3)))
        sp = "%s.stage%d.0" % (bb, s)
        for b in range(s):
            block("%s.branches.%d.0" % (sp, b), ch[b], ch[b])
        for i in range(s):
            for j in range(s):
                if j > i:
                    q = "%s.fuse_layers.%d.%d" % (sp, i, j)
                    gn(q + ".0", ch[j])
                    spec.append((q + ".1.weight", (ch[i], ch[j], 1, 1, 1)))
                elif j < i:
                    for k in range(i - j):
                        q = "%s.fuse_layers.%d.%d.%d" % (sp, i, j, k)
                        gn(q + ".0", ch[j])
                        co = ch[i] if k == i - j - 1 else ch[j]
                        spec.append((q + ".1.weight", (co, ch[j], 3, 3, 3)))
    if c["final_in"] != c["final_out"]:
        spec.append(("backbone.final_conv.weight", (c["final_out"], c["final_in"], 1, 1, 1)))
        spec.append(("backbone.final_conv.bias", (c["final_out"],)))
    if c["head_in"] != c["share"]:
        gn("pose_head.shared_conv.0", c["head_in"])
        spec.append(("pose_head.shared_conv.1.weight", (c["share"], c["head_in"], 3, 3, 3)))
    for h, k in (("reg", c["reg"]), ("hm", c["ncls"])):
        q = "pose_head.tasks.0." + h
        spec.append((q + ".0.weight", (32, c["share"], 3, 3, 3)))
        spec.append((q + ".0.bias", (32,)))
        spec.append((q + ".2.weight", (k, 32, 3, 3, 3)))
        spec.append((q + ".2.bias", (k,)))
    return spec


def synth_state_dict(cfg, seed=0):
    """Deterministic weights from (key, shape) only — regenerable anywhere without shipping 9 MB fixtures.
    Statistics mimic the reference's init (kaiming-uniform-like conv weights, GN gamma~1, beta~0, hm bias -2.19).
    """
    import zlib

    sd = {}
    for key, shape in state_dict_spec(cfg):
        rs = np.random.RandomState((zlib.crc32(key.encode()) + seed) & 0x7FFFFFFF)
        if len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            w = rs.uniform(-1, 1, size=shape) * math.sqrt(1.0 / fan_in) * 1.7
        elif key.endswith("hm.2.bias"):
            w = np.full(shape, -2.19) + rs.uniform(-0.05, 0.05, size=shape)
        elif key.endswith(".bias"):
            w = rs.uniform(-0.1, 0.1, size=shape)
        else:  # GroupNorm gamma
            w = 1.0 + rs.uniform(-0.2, 0.2, size=shape)
        sd[key] = torch.from_numpy(w.astype(np.float32))
    return sd


def reference_init_state_dict(cfg, seed=0):
    """Random-init weights with the distributions the reference's constructors draw from:
    Conv3d default = kaiming_uniform_(a=sqrt(5)) -> U(+-1/sqrt(fan_in)), conv bias U(+-1/sqrt(fan_in)) (torch
    nn.Conv3d.reset_parameters, called at common.py:40,114; hr3d.py:84-185; hrnet3d.py:20; center_head.py:86-91);
    GroupNorm gamma 1 / beta 0; every conv of a non-'hm' head kaiming_init (normal, fan_out, relu, bias 0:
    center_head.py:96-99, torchie/cnn/weight_init.py:32-45); 'hm' last-conv bias = -2.19 (center_head.py:94-95).
    Same distributions as the reference model under a seed, not the same bits (the draw order of its constructors is
    not reproduced); tests/test_oracle_golden.py pins the per-tensor statistics against the reference's own model."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}
    for key, shape in state_dict_spec(cfg):
        if len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            if ".reg." in key:
                fan_out = shape[0] * shape[2] * shape[3] * shape[4]
                w = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
            else:
                b = 1.0 / math.sqrt(fan_in)
                w = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif key.endswith("hm.2.bias"):
            w = torch.full(shape, -2.19)
        elif ".reg." in key and key.endswith(".bias"):
            w = torch.zeros(shape)
        elif key.endswith(".bias") and ("groupnorm" in key or key.split(".")[-2] in ("0",) and "tasks" not in key and "conv1" not in key):
            w = torch.zeros(shape)  # GroupNorm beta
        elif key.endswith(".bias"):
            fan_in = dict(state_dict_spec(cfg))[key[:-4] + "weight"]
            fan_in = fan_in[1] * fan_in[2] * fan_in[3] * fan_in[4]
            b = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:
            w = torch.ones(shape)  # GroupNorm gamma
        sd[key] = w.float()
    return sd


def synth_pose(rs, grid_zyx, voxel=VOXEL_SIZE, pc_range=PC_RANGE):
    """Random 15-joint skeleton: pelvis uniform in the ROI shrunk by 0.5 m (or 25 % of a small
    test grid), joints = pelvis + N(0, 0.3 m) clipped to the ROI."""
    Z, Y, X = grid_zyx
    lo = np.array(pc_range, dtype=np.float64)
    ext = np.array([X * voxel[0], Y * voxel[1], Z * voxel[2]])
    shrink = np.minimum(0.5, 0.25 * ext)
    pelvis = lo + shrink + rs.uniform(0, 1, 3) * (ext - 2 * shrink)
    joints = pelvis[None] + rs.normal(0, 0.3, (15, 3))
    joints[0] = pelvis
    eps = 1e-3
    return np.clip(joints, lo + eps, lo + ext - eps)


def batch_targets(poses, grid_zyx, one_hm):
    # min_radius: 2 for the one_hm configs (hr3d_one_hm.py:107), 1 for hr3d (hr3d.py:108)
    ts = [assign_targets(p, grid_zyx, one_hm, radius=2 if one_hm else 1) for p in poses]
    return {k: torch.from_numpy(np.stack([t[k] for t in ts])) for k in ts[0]}


# ------------------------------------------------------------------------------------------------ evaluation
def abs_pjpe(pred, gt):
    """eval_util.py:10-11 — per-joint Euclidean error, float64 [J]."""
    d = np.asarray(pred, dtype=np.float64) - np.asarray(gt, dtype=np.float64)
    return np.sqrt((d * d).sum(axis=-1))


def pjpe(pred, gt):
    """eval_util.py:5-8 — the same after subtracting joint 0 (the root) on both sides."""
    pred, gt = np.array(pred, dtype=np.float64), np.array(gt, dtype=np.float64)
    return abs_pjpe(pred - pred[:1], gt - gt[:1])


def evaluation(detections, gt, seq_id_to_name):
    """det3d/datasets/cruw_pose/cruw_pose.py:277-310: per-sequence, per-joint mean errors in millimetres, their means,
    and the mean over sequences.  detections: {'seq/frame/rdr_frame': {'keypoints': [(label, x, y, z, score)] * 15}};
    gt: {seq: {frame: [{'pose': [[x, y, z]] * 15}]}}."""
    rel, ab = {}, {}
    for key, val in detections.items():
        seq, frame, _ = key.split("/")
        kp = [p[1:4] for p in val["keypoints"]]
        rel.setdefault(seq, []).append(pjpe(kp, gt[seq][frame][0]["pose"]))
        ab.setdefault(seq, []).append(abs_pjpe(kp, gt[seq][frame][0]["pose"]))
    seq_res = {}
    for seq in rel:
        r, a = np.mean(np.array(rel[seq]), axis=0) * 1000, np.mean(np.array(ab[seq]), axis=0) * 1000
        out = {"MPJPE": np.mean(r), "ABS_MPJPE": np.mean(a)}
        for j in range(r.shape[0]):
            out["PJPE_%d" % j], out["ABS_PJPE_%d" % j] = r[j], a[j]
        seq_res[seq_id_to_name[seq]] = out
    total = {k: np.mean([v[k] for v in seq_res.values()]) for k in ["MPJPE", "ABS_MPJPE"] +
             [p % j for j in range(15) for p in ("PJPE_%d", "ABS_PJPE_%d")]}
    seq_res["ALL"] = total
    return {"results": total, "seq_results": seq_res}


# ------------------------------------------------------------------------------------------------ optimizer step
def adam_step_flat(p, g, m, v, step, lr, mom, beta2=0.99, eps=1e-8, wd=0.01, max_norm=35.0):
    """One training-step update on flat float32 numpy buffers (in place on p, m, v; returns the pre-clip gradient norm):
    OptimizerHook.clip_grads (hooks/optimizer.py:9-12: clip_grad_norm_, coefficient max_norm / (norm + 1e-6) capped at 1),
    OptimWrapper.step's decoupled decay p *= 1 - wd*lr (fastai_optim.py:158-174, true_wd, bn_wd) and
    torch.optim.Adam(betas=(mom, beta2), eps).step() with bias correction at time step `step` >= 1.
    This is the arithmetic rtp_adam_step fuses into one pass (csrc/train_aux.cu)."""
    f = np.float32
    norm = f(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
    coef = f(min(1.0, float(f(max_norm) / (norm + f(1e-6))))) if max_norm > 0 else f(1.0)
    gi = g * coef
    p *= f(1.0) - f(wd) * f(lr)
    m[:] = f(mom) * m + (f(1.0) - f(mom)) * gi
    v[:] = f(beta2) * v + (f(1.0) - f(beta2)) * gi * gi
    bias1 = f(1.0 - mom ** step)
    bias2_sqrt = f(np.sqrt(1.0 - beta2 ** step))
    p -= (f(lr) / bias1) * m / (np.sqrt(v) / bias2_sqrt + f(eps))
    return float(norm)


# ------------------------------------------------------------------------------------------------ space-to-depth identity
def s2d_view(x):
    """[N,C,Z,Y,X] (even extents) -> [N,8C,Z/2,Y/2,X/2], channel (pz*4 + px*2 + py)*C + c = x[n, c, 2z+pz, 2y+py, 2x+px]
    (px is the parity of the LAST tensor axis, py of the one before: the P8 x / y axes)."""
    parts = [x[:, :, pz::2, py::2, px::2] for pz in (0, 1) for px in (0, 1) for py in (0, 1)]
    return torch.cat(parts, dim=1)


def s2d_expand_weight(w):
    """[Cout,Cin,3,3,3] of a stride-2 pad-1 conv -> [Cout,8Cin,3,3,3] of the equivalent stride-1 pad-1 conv over s2d_view:
    input index i = 2o + k - 1, so per axis tap k=0 reads parity 1 at offset -1 (view tap 0), k=1 parity 0 at offset 0
    (view tap 1), k=2 parity 1 at offset 0 (view tap 1); every other (parity, view tap) pair is zero.  Restates
    weight_s2d_expand_kernel (csrc/layout.cu)."""
    Cout, Cin = w.shape[:2]
    we = torch.zeros((Cout, 8 * Cin, 3, 3, 3), dtype=w.dtype)
    pt = {0: (1, 0), 1: (0, 1), 2: (1, 1)}  # k -> (parity, view tap)
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                (pz, tz), (py, ty), (px, tx) = pt[kz], pt[ky], pt[kx]
                par = pz * 4 + px * 2 + py
                we[:, par * Cin:(par + 1) * Cin, tz, ty, tx] = w[:, :, kz, ky, kx]
    return we
