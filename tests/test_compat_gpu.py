"""GPU tests of the drop-in boundary: the det3d-style modules (build_detector -> RadarPoseNet / HRNet3D / CenterHead)
driven exactly as the reference's trainer / test tool drive them (trainer.py:370-396, hooks/optimizer.py:14-24,
tools/test.py:203-206), checked against the CPU oracle with the tolerances of test_engine_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O
from oracle import make_golden as G

pytestmark = pytest.mark.gpu

POSE = ["Pelvis", "Right_Hip", "Right_Knee", "Right_Ankle", "Left_Hip", "Left_Knee", "Left_Ankle", "Thomx", "Head",
        "Left_Shoulder", "Left_Elbow", "Left_Wrist", "Right_Shoulder", "Right_Elbow", "Right_Wrist"]


class TestCfg(dict):
    __getattr__ = dict.__getitem__


def build(cfg):
    from rtpose_b200 import det3d_compat as D
    c = O.CONFIGS[cfg]
    model_cfg = dict(type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
                     backbone=dict(type="HRNet3D", backbone_cfg=c["arch"], final_conv_in=c["final_in"], final_conv_out=c["final_out"],
                                   final_fuse=c["fuse"], ds_factor=1),
                     pose_head=dict(type="CenterHead", tasks=[dict(num_class=c["ncls"], class_names=POSE[:c["ncls"]])],
                                    in_channels=c["head_in"], share_conv_channel=c["share"], dataset="cruw_pose", weight=c["weight"],
                                    code_weights=c["code_weights"], common_heads={"reg": (c["reg"], 2)}, dcn_head=False),
                     neck=None)
    test_cfg = TestCfg(post_center_limit_range=[], score_threshold=0.0, pc_range=list(O.PC_RANGE), out_size_factor=[1, 1, 1],
                       voxel_size=list(O.VOXEL_SIZE))
    m = D.build_detector(model_cfg, train_cfg=None, test_cfg=test_cfg)
    m.load_state_dict(O.synth_state_dict(cfg), strict=True)
    return m.cuda(), test_cfg


def example_of(x, tgt, batch):
    return {"rdr": {"rdr_tensor": torch.from_numpy(x).cuda(), "hm": [tgt["hm"].cuda()], "anno_pose": [tgt["anno_pose"].cuda()],
                    "ind": [tgt["ind"].cuda()], "mask": [tgt["mask"].cuda()], "cat": [tgt["cat"].cuda()]},
            "meta": [{"frame": i} for i in range(batch)]}


@pytest.mark.parametrize("cfg", ["hr3d_one_hm_doppler", "hr3d"])
def test_detector_train_step_and_inference(cfg):
    from test_engine_gpu import grad_report, oracle_run
    grid, batch = (8, 16, 24), 2
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=21)
    model, test_cfg = build(cfg)
    model.train()
    losses = model(example_of(x, tgt, batch), return_loss=True)
    assert set(losses) == {"loss", "hm_loss", "loc_loss", "loc_loss_elem", "num_positive"}
    assert not losses["hm_loss"][0].is_cuda and not losses["loc_loss_elem"][0].is_cuda  # CPU copies, like the reference
    loss = sum(losses["loss"])  # trainer.py:74-89
    model.zero_grad()
    (loss * 2.0).backward()     # an upstream scale factor (e.g. a GradScaler) must flow through linearly
    r_hm, r_reg, r_loss, r_grads = oracle_run(x, O.synth_state_dict(cfg), cfg, tgt, False)
    b_hm, b_reg, b_loss, b_grads = oracle_run(x, O.synth_state_dict(cfg), cfg, tgt, True)
    assert abs(float(loss) - r_loss) <= 1.5e-2 * abs(r_loss)
    got = {k: p.grad.cpu() / 2.0 for k, p in model.named_parameters() if p.grad is not None}
    rel, cos = grad_report(got, r_grads)
    brel, bcos = grad_report(b_grads, r_grads)
    assert cos >= min(0.97, bcos - 0.01), (cos, bcos)
    assert set(got) >= set(r_grads)
    # second step with changed weights must repack them (the pack cache keys on the parameter version)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.5)
    l2 = sum(model(example_of(x, tgt, batch), return_loss=True)["loss"])
    assert abs(float(l2) - float(loss)) > 1e-3 * abs(float(loss))
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)
    # inference path: list of {'keypoints': [(label, x, y, z, score)...], 'metadata': meta}
    model.eval()
    with torch.no_grad():
        dets = model(example_of(x, tgt, batch), return_loss=False)
    assert len(dets) == batch and dets[1]["metadata"] == {"frame": 1}
    preds, _ = model.pose_head(model.extract_feat({"rdr_tensor": torch.from_numpy(x).cuda()}))
    hm, reg = preds[0]["hm"].float().cpu(), preds[0]["reg"].float().cpu()
    kps, _ = O.decode(hm, reg)
    for n in range(batch):
        assert len(dets[n]["keypoints"]) == len(kps[n])
        a, b = np.array(dets[n]["keypoints"], dtype=np.float64), np.array(kps[n], dtype=np.float64)
        np.testing.assert_array_equal(a[:, 0], b[:, 0])
        np.testing.assert_allclose(a[:, 1:], b[:, 1:], rtol=1e-5, atol=1e-5)


def test_standalone_backbone_and_head_modules_autograd():
    """HRNet3D.forward and CenterHead.forward/loss/predict used on their own (NCDHW fp32 tensors in and out)."""
    from test_engine_gpu import oracle_run
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 16), 1
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=22)
    model, test_cfg = build(cfg)
    xt = torch.from_numpy(x).cuda()
    feats = model.backbone(xt)
    assert feats.shape == (batch, 128, *grid) and feats.dtype == torch.float32 and feats.requires_grad
    preds, same = model.pose_head(feats)
    assert same is feats and preds[0]["hm"].shape == (batch, 1, *grid) and preds[0]["reg"].shape == (batch, 45, *grid)
    ex = example_of(x, tgt, batch)["rdr"]
    hm_raw = preds[0]["hm"].detach().clone()
    losses = model.pose_head.loss(ex, preds, test_cfg)
    assert float(preds[0]["hm"].min()) >= 1e-4  # loss() leaves the clamped sigmoid in place, like the reference
    model.zero_grad()
    losses["loss"][0].backward()
    r_hm, r_reg, r_loss, r_grads = oracle_run(x, O.synth_state_dict(cfg), cfg, tgt, False)
    assert abs(float(losses["loss"][0]) - r_loss) <= 1.5e-2 * abs(r_loss)
    g = model.backbone.backbone.layer1.conv2.conv.weight.grad
    r = r_grads["backbone.backbone.layer1.conv2.conv.weight"]
    cos = float((g.cpu().flatten().double() @ r.flatten().double()) / (g.norm().item() * r.norm().item()))
    assert cos > 0.9, cos
    dets = model.pose_head.predict({"meta": [{}]}, [{"hm": hm_raw, "reg": preds[0]["reg"].detach()}], test_cfg)
    assert len(dets) == 1 and len(dets[0]["keypoints"]) == 15


def test_graphed_train_step_equals_eager_and_follows_weight_updates():
    """RadarPoseNet.cuda_graph = True: same losses and gradients as the eager path, for new inputs and after an
    optimizer update of the weights (the captured graph repacks the weights on every replay)."""
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 24), 2
    eager, _ = build(cfg)
    graphed, _ = build(cfg)
    graphed.cuda_graph = True
    eager.train(); graphed.train()
    for m in (eager, graphed):
        m.pose_head.sync_free_losses = True
    opt_e = torch.optim.SGD(eager.parameters(), lr=1e-3)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=1e-3)
    for it in range(3):
        x, poses, tgt = G.make_example(cfg, batch, grid, seed=40 + it)
        out = []
        for m, opt in ((eager, opt_e), (graphed, opt_g)):
            opt.zero_grad(set_to_none=True)
            losses = m(example_of(x, tgt, batch), return_loss=True)
            (sum(losses["loss"]) * 0.5).backward()
            out.append((float(losses["loss"][0]), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
            opt.step()
        (le, ge), (lg, gg) = out
        assert le == lg, (it, le, lg)
        assert set(ge) == set(gg)
        for k in ge:
            assert torch.equal(ge[k], gg[k]), (it, k)
    assert graphed._graph_state.get("graph") is not None


def test_second_forward_before_backward_raises_instead_of_wrong_gradients():
    """One engine = one tape: backward of a forward whose tape a later forward recycled must fail loudly (eager and
    graphed), not return the other forward's gradients."""
    from rtpose_b200.lib import RtpError
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 16), 1
    for graphed in (False, True):
        model, _ = build(cfg)
        model.cuda_graph = graphed
        model.pose_head.sync_free_losses = True
        model.train()
        xa, _, ta = G.make_example(cfg, batch, grid, seed=61)
        xb, _, tb = G.make_example(cfg, batch, grid, seed=62)
        la = model(example_of(xa, ta, batch), return_loss=True)["loss"][0]
        lb = model(example_of(xb, tb, batch), return_loss=True)["loss"][0]
        with pytest.raises(RtpError):
            la.backward()
        model.zero_grad(set_to_none=True)
        lb2 = model(example_of(xb, tb, batch), return_loss=True)["loss"][0]
        lb2.backward()  # the latest forward is still fine
        assert model.backbone.backbone.layer1.conv2.conv.weight.grad is not None


def test_flat_adam_step_invalidates_the_weight_packs_of_an_eager_engine():
    """FlatAdam rewrites the parameters through raw pointers; the engine's bf16 weight packs are keyed on the parameters'
    version counters, so the optimizer must bump them — otherwise an eager engine keeps convolving with step-0 weights."""
    from rtpose_b200.engine import Engine
    from rtpose_b200.optim import FlatAdam
    from rtpose_b200.p8 import P8
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 16), 1
    c = O.CONFIGS[cfg]
    sd = O.synth_state_dict(cfg)
    total = sum(v.numel() for v in sd.values())
    flat = torch.empty(total, dtype=torch.float32, device="cuda")
    gflat = torch.zeros_like(flat)
    params, grads, o = {}, {}, 0
    for k, v in sd.items():
        params[k] = flat[o:o + v.numel()].view(v.shape)
        params[k].copy_(v)
        grads[k] = gflat[o:o + v.numel()].view(v.shape)
        o += v.numel()
    eng = Engine(c["arch"], c["fuse"], params, c["reg"], c["ncls"], c["weight"], c["code_weights"])
    opt = FlatAdam(flat, gflat, wd=0.01, max_norm=35.0)
    x, _, tgt = G.make_example(cfg, batch, grid, seed=63)
    xp = P8.from_ncdhw(torch.from_numpy(x).cuda())
    tg = [tgt[k].cuda() for k in ("hm", "ind", "mask", "cat", "anno_pose")]
    losses = []
    for it in range(3):
        hm, reg = eng.forward(xp, True)
        losses.append(float(eng.loss(hm, reg, *tg)[0]))
        eng.backward(grads)
        opt.step(5e-3, 0.9)
    assert losses[1] != losses[0] and losses[2] != losses[1], losses
    # the same three steps with explicitly invalidated packs give the same losses: the version bump alone is enough
    for k, v in sd.items():
        params[k].copy_(v)
    torch.autograd.graph.increment_version(flat)
    opt2 = FlatAdam(flat, gflat, wd=0.01, max_norm=35.0)
    ref = []
    for it in range(3):
        eng.packs.invalidate()
        hm, reg = eng.forward(xp, True)
        ref.append(float(eng.loss(hm, reg, *tg)[0]))
        eng.backward(grads)
        opt2.step(5e-3, 0.9)
    assert ref == losses, (ref, losses)
