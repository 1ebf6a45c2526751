"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by executing the reference's own unmodified modules
(imported from /root/reference through oracle/ref_loader.py) on seeded synthetic inputs.

    python oracle/make_golden.py            # writes tests/golden/<cfg>_g<Z>x<Y>x<X>.npz

Weights come from `hrpose_oracle.synth_state_dict` (a pure function of parameter name + shape), loaded into the
reference model through `load_state_dict(strict=True)` — so the fixture only has to carry inputs and outputs.
The reference model is fully convolutional, so a reduced grid (8x16x24) exercises exactly the code the
16x64x160 grid does; the file stays < 1 MB per config.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import hrpose_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

POSE_NAMES = ["Pelvis", "Right_Hip", "Right_Knee", "Right_Ankle", "Left_Hip", "Left_Knee", "Left_Ankle", "Thomx",
              "Head", "Left_Shoulder", "Left_Elbow", "Left_Wrist", "Right_Shoulder", "Right_Elbow", "Right_Wrist"]


def ref_model_cfg(cfg):
    """The `model=` dict of configs/cruw_pose/<cfg>.py (values transcribed; structure required by the builder)."""
    c = O.CONFIGS[cfg]
    tasks = [dict(num_class=c["ncls"], class_names=POSE_NAMES[:c["ncls"]])]
    return dict(
        type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
        backbone=dict(type="HRNet3D", backbone_cfg=c["arch"], final_conv_in=c["final_in"],
                      final_conv_out=c["final_out"], final_fuse=c["fuse"], ds_factor=1),
        pose_head=dict(type="CenterHead", tasks=tasks, in_channels=c["head_in"], share_conv_channel=c["share"],
                       dataset="cruw_pose", weight=c["weight"], code_weights=c["code_weights"],
                       common_heads={"reg": (c["reg"], 2)}, dcn_head=False),
        neck=None)


def ref_test_cfg():
    return ref_loader.AttrDict(post_center_limit_range=[], score_threshold=0.0, pc_range=list(O.PC_RANGE),
                               out_size_factor=[1, 1, 1], voxel_size=list(O.VOXEL_SIZE))


def build_reference(cfg):
    mods = ref_loader.load()
    model = mods["build_detector"](ref_model_cfg(cfg), train_cfg=None, test_cfg=ref_test_cfg())
    sd = O.synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    return model, sd


def synth_input(cfg, batch, grid, seed):
    c = O.CONFIGS[cfg]
    rs = np.random.RandomState(seed)
    u = rs.uniform(-0.2, 1.0, size=(batch, c["in_ch"]) + tuple(grid)).astype(np.float32)
    return np.maximum(u, 0.0)


def make_example(cfg, batch, grid, seed):
    c = O.CONFIGS[cfg]
    x = synth_input(cfg, batch, grid, seed)
    rs = np.random.RandomState(seed + 1)
    poses = [O.synth_pose(rs, grid) for _ in range(batch)]
    tgt = O.batch_targets(poses, grid, one_hm=(c["ncls"] == 1))
    return x, poses, tgt


def run_reference(cfg, batch, grid, seed):
    model, sd = build_reference(cfg)
    x, poses, tgt = make_example(cfg, batch, grid, seed)
    xt = torch.from_numpy(x)
    example = {"rdr": {"rdr_tensor": xt, "hm": [tgt["hm"]], "anno_pose": [tgt["anno_pose"]], "ind": [tgt["ind"]],
                       "mask": [tgt["mask"]], "cat": [tgt["cat"]]}, "meta": [{"i": i} for i in range(batch)]}
    model.train()
    feats = model.extract_feat({"rdr_tensor": xt})
    preds, _ = model.pose_head(feats)
    hm = preds[0]["hm"].detach().clone()
    reg = preds[0]["reg"].detach().clone()
    losses = model(example, return_loss=True)
    loss = losses["loss"][0]
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.eval()
    with torch.no_grad():
        dets = model(example, return_loss=False)
    out = dict(x_checksum=np.float64(x.astype(np.float64).sum()), poses=np.stack(poses), hm=hm.numpy(), reg=reg.numpy(),
               feats=feats.detach().numpy()[:, :4],
               loss=np.float32(loss.item()), hm_loss=np.float32(losses["hm_loss"][0].item()),
               loc_loss=np.float32(losses["loc_loss"][0].item()), loc_loss_elem=losses["loc_loss_elem"][0].numpy(),
               num_positive=np.float32(losses["num_positive"][0].item()))
    kp = np.zeros((batch, 15, 5), np.float64)
    kpn = np.zeros((batch,), np.int64)
    for n, d in enumerate(dets):
        kpn[n] = len(d["keypoints"])
        for j, t in enumerate(d["keypoints"]):
            kp[n, j] = np.array(t, dtype=np.float64)
    out["keypoints"] = kp
    out["num_keypoints"] = kpn
    # gradient fixtures: global norm per parameter + a few full tensors
    names = sorted(grads)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(grads[k].norm()) for k in names], np.float64)
    for k in ("backbone.backbone.layer1.conv2.conv.weight", "pose_head.tasks.0.hm.2.weight",
              "backbone.backbone.stage3.0.fuse_layers.0.2.1.weight",
              "backbone.backbone.stage4.0.branches.3.0.conv3.groupnorm.weight"):
        if k in grads:
            out["grad::" + k] = grads[k].numpy()
    for k in ("hm", "ind", "mask", "cat", "anno_pose"):
        out["tgt_" + k] = tgt[k].numpy()
    return out


GOLDEN = [("hr3d", 2, (8, 16, 24), 11), ("hr3d_one_hm_doppler", 2, (8, 16, 24), 12), ("hr3d_one_hm", 1, (8, 16, 16), 13)]


def main():
    outdir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for cfg, batch, grid, seed in GOLDEN:
        out = run_reference(cfg, batch, grid, seed)
        out["meta"] = np.array([cfg, str(batch), "x".join(map(str, grid)), str(seed)])
        path = os.path.join(outdir, "%s_g%s.npz" % (cfg, "x".join(map(str, grid))))
        np.savez_compressed(path, **out)
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss", float(out["loss"]))


if __name__ == "__main__":
    main()
