"""GPU parity tests, path level: the whole HRRadarPose forward / loss / backward / decode through the C-ABI
kernels against (a) the golden vectors produced by the reference's own modules (tests/golden) and (b) the CPU
oracle on fresh seeded inputs.

Tolerances (SURVEY.md §8c contract (3)): the yardstick is the oracle's OWN bf16-autocast vs fp32 spread measured
in the same test on the same inputs and weights (torch.autocast on CPU runs the reference's arithmetic in bf16).
We require: RMS error of hm and reg <= 1.5 x the autocast RMS error, max|d hm| and max|d reg| <= 2 x the autocast
maximum (a maximum over ~10^4 voxels is a noisy statistic; floor 2 % of the tensor's std), loss within 1.5 %,
global gradient cosine >= min(0.97, the autocast cosine - 0.01), per-parameter gradient rel-L2 median <= 1.5 x the
autocast median.  Decoded indices are bit-exact at the decode boundary (same heatmap in -> same index out).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O
from oracle import make_golden as G

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_g*x*x*.npz")))  # model goldens (make_golden.py)


def build_engine(cfg, sd=None):
    from rtpose_b200.engine import Engine
    c = O.CONFIGS[cfg]
    sd = sd if sd is not None else O.synth_state_dict(cfg)
    params = {k: v.cuda() for k, v in sd.items()}
    return Engine(c["arch"], c["fuse"], params, c["reg"], c["ncls"], c["weight"], c["code_weights"]), params


def run_engine(eng, params, x, tgt, train=True):
    from rtpose_b200.p8 import P8
    xp = P8.from_ncdhw(torch.from_numpy(x).cuda())
    hm, reg = eng.forward(xp, train)
    out = {"hm": hm.to_ncdhw().cpu(), "reg": reg.to_ncdhw().cpu()}
    if tgt is not None:
        loss = eng.loss(hm, reg, tgt["hm"].cuda(), tgt["ind"].cuda(), tgt["mask"].cuda(), tgt["cat"].cuda(),
                        tgt["anno_pose"].cuda(), with_grad=train)
        out["loss"] = loss.cpu()
        if train:
            grads = {k: torch.zeros_like(v) for k, v in params.items()}
            touched = eng.backward(grads)
            out["grads"] = {k: grads[k].cpu() for k in touched}
    out["decode"] = tuple(t.cpu() for t in eng.decode(hm, reg, O.VOXEL_SIZE, O.PC_RANGE))
    torch.cuda.synchronize()
    return out, hm, reg


def oracle_run(x, sd, cfg, tgt, autocast):
    """fp32 (or bf16-autocast) CPU oracle: returns hm, reg, loss, grads."""
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    c = O.CONFIGS[cfg] if isinstance(cfg, str) else cfg
    torch.set_num_threads(8)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        preds = O.forward(torch.from_numpy(x), sdr, cfg)
    preds = {k: v.float() for k, v in preds.items()}
    L = O.head_loss(preds, tgt, c["weight"], c["code_weights"])
    L["loss"].backward()
    grads = {k: v.grad for k, v in sdr.items() if v.grad is not None and float(v.grad.norm()) > 0}
    return preds["hm"].detach(), preds["reg"].detach(), float(L["loss"]), grads


def grad_report(got, ref):
    names = sorted(ref)
    rel, dots, n1, n2 = [], 0.0, 0.0, 0.0
    for k in names:
        r = ref[k].double().flatten()
        g = got[k].double().flatten() if k in got else torch.zeros_like(r)
        if float(r.norm()) > 0:
            rel.append(float((g - r).norm() / r.norm()))
        dots += float(g @ r)
        n1 += float(g @ g)
        n2 += float(r @ r)
    return np.array(rel), dots / max(np.sqrt(n1 * n2), 1e-30)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_against_reference_golden(path):
    g = np.load(path, allow_pickle=False)
    cfg, batch, grid, seed = [str(v) for v in g["meta"]]
    batch, grid, seed = int(batch), tuple(int(v) for v in grid.split("x")), int(seed)
    x, poses, tgt = G.make_example(cfg, batch, grid, seed)
    eng, params = build_engine(cfg)
    out, hm, reg = run_engine(eng, params, x, tgt)
    ref_hm, ref_reg = torch.from_numpy(g["hm"]), torch.from_numpy(g["reg"])
    e_hm = (out["hm"] - ref_hm).abs().max().item()
    e_reg = (out["reg"] - ref_reg).abs().max().item()
    # yardstick: the oracle's bf16-autocast spread against the same golden tensors
    b_hm, b_reg, _, b_grads = oracle_run(x, O.synth_state_dict(cfg), cfg, tgt, True)
    s_hm = max((b_hm - ref_hm).abs().max().item(), 0.02 * ref_hm.std().item())
    s_reg = max((b_reg - ref_reg).abs().max().item(), 0.02 * ref_reg.std().item())
    print("hm max err %.4g (autocast spread %.4g, std %.3g), reg max err %.4g (spread %.4g, std %.3g)" %
          (e_hm, s_hm, ref_hm.std(), e_reg, s_reg, ref_reg.std()))
    assert e_hm <= 2.0 * s_hm, (e_hm, s_hm)
    assert e_reg <= 2.0 * s_reg, (e_reg, s_reg)
    for ours, auto, ref_t, name in ((out["hm"], b_hm, ref_hm, "hm"), (out["reg"], b_reg, ref_reg, "reg")):
        rms_o, rms_a = (ours - ref_t).pow(2).mean().sqrt().item(), (auto - ref_t).pow(2).mean().sqrt().item()
        assert rms_o <= 1.5 * max(rms_a, 0.005 * ref_t.std().item()), (name, rms_o, rms_a)
    loss = out["loss"][0].item()
    assert abs(loss - float(g["loss"])) <= 1.5e-2 * abs(float(g["loss"])), (loss, float(g["loss"]))
    assert out["loss"][3].item() == float(g["num_positive"])
    # gradients: norms of every parameter + a few full tensors
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    rel = []
    for k, r in norms.items():
        if r > 0:
            assert k in out["grads"], "no gradient produced for %s" % k
            rel.append(abs(float(out["grads"][k].norm()) - r) / r)
    print("per-parameter grad-norm rel err: median %.3g max %.3g" % (np.median(rel), np.max(rel)))
    assert np.median(rel) <= 0.1
    for k in g.files:
        if k.startswith("grad::"):
            r = torch.from_numpy(g[k]).double().flatten()
            q = out["grads"][k[6:]].double().flatten()
            cos = float(q @ r / (q.norm() * r.norm() + 1e-30))
            a = b_grads[k[6:]].double().flatten()  # the oracle's own bf16-autocast gradient of the same tensor
            acos = float(a @ r / (a.norm() * r.norm() + 1e-30))
            print("%s: cosine %.4f (autocast %.4f)" % (k, cos, acos))
            assert cos >= min(0.98, acos - 0.03), (k, cos, acos)


@pytest.mark.parametrize("cfg,grid,batch", [("hr3d_one_hm_doppler", (8, 16, 24), 2), ("hr3d", (8, 16, 16), 1),
                                            ("hr3d_one_hm_doppler_phase", (8, 16, 16), 1),
                                            # ragged grid: every resolution level has an odd extent somewhere (6,10,14 ->
                                            # 3,5,7 -> 2,3,4 -> 1,2,2), so no stride-2 conv can use the s2d view and the
                                            # upsample factors are not powers of two
                                            ("hr3d_one_hm_doppler", (6, 10, 14), 3)])
def test_against_oracle_full_gradient(cfg, grid, batch):
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=101)
    sd = O.synth_state_dict(cfg, seed=3)
    eng, params = build_engine(cfg, sd)
    out, hm, reg = run_engine(eng, params, x, tgt)
    r_hm, r_reg, r_loss, r_grads = oracle_run(x, sd, cfg, tgt, False)
    b_hm, b_reg, b_loss, b_grads = oracle_run(x, sd, cfg, tgt, True)
    s_hm = max((b_hm - r_hm).abs().max().item(), 0.02 * r_hm.std().item())
    s_reg = max((b_reg - r_reg).abs().max().item(), 0.02 * r_reg.std().item())
    e_hm = (out["hm"] - r_hm).abs().max().item()
    e_reg = (out["reg"] - r_reg).abs().max().item()
    print("hm err %.4g (spread %.4g) reg err %.4g (spread %.4g) loss %.5g vs %.5g" %
          (e_hm, s_hm, e_reg, s_reg, out["loss"][0].item(), r_loss))
    assert e_hm <= 2.0 * s_hm and e_reg <= 2.0 * s_reg
    for ours, auto, ref_t, name in ((out["hm"], b_hm, r_hm, "hm"), (out["reg"], b_reg, r_reg, "reg")):
        rms_o, rms_a = (ours - ref_t).pow(2).mean().sqrt().item(), (auto - ref_t).pow(2).mean().sqrt().item()
        print("%s rms err %.4g (autocast %.4g)" % (name, rms_o, rms_a))
        assert rms_o <= 1.5 * max(rms_a, 0.005 * ref_t.std().item()), (name, rms_o, rms_a)
    assert abs(out["loss"][0].item() - r_loss) <= 1.5e-2 * abs(r_loss)
    rel, cos = grad_report(out["grads"], r_grads)
    brel, bcos = grad_report(b_grads, r_grads)
    print("grad: global cosine %.5f (autocast %.5f), per-param rel-L2 median %.3g (autocast %.3g) max %.3g" %
          (cos, bcos, np.median(rel), np.median(brel), rel.max()))
    assert cos >= min(0.97, bcos - 0.01)
    assert np.median(rel) <= 1.5 * np.median(brel)
    assert set(out["grads"]) >= set(r_grads), "parameters without a gradient: %s" % (set(r_grads) - set(out["grads"]))
    # decode bit-exact at the decode boundary: oracle decode of OUR heatmap == our decode
    kps, ref_idx = O.decode(out["hm"], out["reg"])
    assert out["decode"][0].tolist() == ref_idx


def test_inference_matches_training_forward():
    cfg, grid = "hr3d_one_hm_doppler", (8, 16, 24)
    x, poses, tgt = G.make_example(cfg, 2, grid, seed=7)
    eng, params = build_engine(cfg)
    a, _, _ = run_engine(eng, params, x, None, train=False)
    b, _, _ = run_engine(eng, params, x, tgt, train=True)
    assert torch.equal(a["hm"], b["hm"]) and torch.equal(a["reg"], b["reg"])


def test_stream_schedules_are_bitwise_equivalent():
    """Every execution schedule of the engine — one stream, branch / fuse / weight-gradient side streams, and the optional
    stream-parallel fuse backward with event-ordered gradient accumulation — produces bit-identical outputs, loss and
    gradients (all reductions have a fixed order; accumulation into shared gradients follows program order)."""
    from rtpose_b200 import ops
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 24), 2
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=33)
    results = []
    for branches, fuse, fuse_bwd, async_wgrad in ((False, False, False, False), (True, True, False, True),
                                                  (True, True, True, True)):
        eng, params = build_engine(cfg)
        eng.parallel_branches, eng.parallel_fuse, eng.parallel_fuse_bwd = branches, fuse, fuse_bwd
        old = ops.ASYNC_WGRAD
        ops.ASYNC_WGRAD = async_wgrad
        try:
            out, _, _ = run_engine(eng, params, x, tgt, True)
        finally:
            ops.ASYNC_WGRAD = old
        results.append(out)
    ref = results[0]
    for out in results[1:]:
        assert torch.equal(out["hm"], ref["hm"]) and torch.equal(out["reg"], ref["reg"])
        assert torch.equal(out["loss"], ref["loss"])
        assert set(out["grads"]) == set(ref["grads"])
        for k in ref["grads"]:
            assert torch.equal(out["grads"][k], ref["grads"][k]), k


@pytest.mark.parametrize("grid,batch", [((8, 16, 24), 2), ((16, 64, 160), 2)], ids=["small", "full_grid"])
def test_regression_branch_evaluated_around_the_targets_only(grid, batch):
    """Engine.forward(reg_targets=ind): in training the loss gathers the regression map at the target voxels only, so the
    regression branch of the head runs on the (sample, tile) units around them (forward AND backward).  Loss and every
    parameter gradient must agree with the dense evaluation; the regression map must agree at the target voxels; switching
    the sparse paths off (RTP_NO_SPARSE_* at import time) is the dense evaluation used as the reference here."""
    from rtpose_b200 import ops
    from rtpose_b200.p8 import P8
    cfg = "hr3d_one_hm_doppler"
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=21)
    eng, params = build_engine(cfg)

    def run(sparse):
        xp = P8.from_ncdhw(torch.from_numpy(x).cuda())
        ind = tgt["ind"].cuda()
        old = (ops.USE_SPARSE_REG, ops.USE_SPARSE_UNITS, ops.USE_SPARSE_FWD)
        if not sparse:
            ops.USE_SPARSE_REG = ops.USE_SPARSE_UNITS = ops.USE_SPARSE_FWD = False
        try:
            hm, reg = eng.forward(xp, True, reg_targets=ind if sparse else None)
            loss = eng.loss(hm, reg, tgt["hm"].cuda(), ind, tgt["mask"].cuda(), tgt["cat"].cuda(), tgt["anno_pose"].cuda())
            grads = {k: torch.zeros_like(v) for k, v in params.items()}
            touched = eng.backward(grads)
        finally:
            ops.USE_SPARSE_REG, ops.USE_SPARSE_UNITS, ops.USE_SPARSE_FWD = old
        torch.cuda.synchronize()
        return hm.to_ncdhw().cpu(), reg.to_ncdhw().cpu(), loss.cpu(), {k: grads[k].cpu() for k in touched}

    hm_d, reg_d, loss_d, g_d = run(False)
    hm_s, reg_s, loss_s, g_s = run(True)
    # (the heat-map half of the merged head conv is its own launch in the sparse evaluation: same operands, another
    # accumulation grouping — equal to fp32 round-off before the bf16 store)
    assert float((hm_s - hm_d).abs().max()) <= 2.0 ** -7 * float(hm_d.abs().max())
    assert float((hm_s != hm_d).float().mean()) < 0.02
    Z, Y, X = grid
    ind = tgt["ind"]
    for n in range(batch):
        for j in range(ind.shape[1]):
            i = int(ind[n, j])
            z, y, xx = i // (Y * X), (i % (Y * X)) // X, i % X
            assert float((reg_s[n, :, z, y, xx] - reg_d[n, :, z, y, xx]).abs().max()) <= 2.0 ** -7 * float(reg_d.abs().max()), (n, j)
    assert torch.allclose(loss_s, loss_d, rtol=2e-3, atol=1e-5), (loss_s, loss_d)
    assert set(g_s) == set(g_d)
    for k in g_d:
        den = float(g_d[k].norm()) + 1e-12
        rel = float((g_s[k] - g_d[k]).norm()) / den
        assert rel <= 2e-2, (k, rel)   # bf16 round-off of the re-grouped head convs propagates through the whole backward pass


def test_gradient_ready_notifications_follow_the_tape():
    """Engine.set_grad_groups: after the first (learning) backward pass every group is reported exactly once per pass, in
    backward order (tail of the flat buffer first), and only after its last gradient has been issued — the hook the
    sliced all-reduce hangs on (dist.SlicedAllReduce)."""
    from rtpose_b200 import dist as rdist
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 16, 24), 2
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=5)
    eng, params = build_engine(cfg)
    total = sum(v.numel() for v in params.values())
    flat = torch.zeros(total, device="cuda")
    sar = rdist.SlicedAllReduce(flat, [(k, v.numel()) for k, v in params.items()], 3, world=1).attach(eng)
    assert [hi for _, hi in sar.spans][0] == total and sar.spans[-1][0] == 0
    assert sum(len(g) for g in sar.groups) == len(params)
    seen = []
    orig = sar.on_ready
    def spy(k):
        seen.append((k, set(eng._touched)))
        orig(k)
    eng._on_ready = spy
    for it in range(3):
        seen.clear()
        out, _, _ = run_engine(eng, params, x, tgt)
        assert sorted(k for k, _ in seen) == [0, 1, 2], seen
        if it > 0:  # learnt: reported in backward order, each after all of its gradients were issued
            assert [k for k, _ in seen] == [0, 1, 2]
            for k, touched in seen:
                assert set(sar.groups[k]) & set(out["grads"]) <= touched, k
            # the first slice is reported before the backward pass has reached the stem
            assert "backbone.backbone.layer1.conv2.conv.weight" not in seen[0][1]


def test_shared_conv_branch_of_the_head():
    """CenterHead with in_channels != share_conv_channel: the GN -> Conv3d(3x3x3, no bias) -> ReLU `shared_conv`
    (center_head.py:203-211) in front of the task heads — not used by the shipped configs, but part of the head's API."""
    cfg = dict(O.CONFIGS["hr3d"], share=64)  # hr3d backbone ('top', 32 channels out) feeding a 64-channel shared conv
    grid, batch = (8, 16, 16), 2
    x, poses, tgt = G.make_example("hr3d", batch, grid, seed=77)
    sd = O.synth_state_dict(cfg, seed=5)
    assert "pose_head.shared_conv.1.weight" in sd and sd["pose_head.tasks.0.reg.0.weight"].shape[1] == 64
    from rtpose_b200.engine import Engine
    params = {k: v.cuda() for k, v in sd.items()}
    eng = Engine(cfg["arch"], cfg["fuse"], params, cfg["reg"], cfg["ncls"], cfg["weight"], cfg["code_weights"])
    out, hm, reg = run_engine(eng, params, x, tgt)
    r_hm, r_reg, r_loss, r_grads = oracle_run(x, sd, cfg, tgt, False)
    b_hm, b_reg, b_loss, b_grads = oracle_run(x, sd, cfg, tgt, True)
    for ours, a, r, name in ((out["hm"], b_hm, r_hm, "hm"), (out["reg"], b_reg, r_reg, "reg")):
        spread = max((a - r).abs().max().item(), 0.02 * r.std().item())
        assert (ours - r).abs().max().item() <= 2.0 * spread, name
    assert abs(out["loss"][0].item() - r_loss) <= 1.5e-2 * abs(r_loss)
    rel, cos = grad_report(out["grads"], r_grads)
    brel, bcos = grad_report(b_grads, r_grads)
    assert cos >= min(0.97, bcos - 0.01), (cos, bcos)
    for k in ("pose_head.shared_conv.0.weight", "pose_head.shared_conv.0.bias", "pose_head.shared_conv.1.weight"):
        assert k in out["grads"] and float(out["grads"][k].abs().sum()) > 0, k


def test_sibling_stride2_convs_share_one_normalised_view(monkeypatch):
    """The stride-2 fuse convs out of branch 0 of an HR module (2 in stage 3, 3 in stage 4) read ONE space-to-depth view of
    xhat with their GroupNorm affine folded into the conv (csrc/s2d_shared.cu): forward, every parameter gradient (incl. the
    folded gamma / beta of those GroupNorms, which come out of the weight gradient) and dL/dx agree with the per-conv path and
    with the fp32 oracle.  synth weights: beta != 0, so the border-class bias is exercised."""
    from rtpose_b200 import lib, ops
    cfg, grid, batch = "hr3d_one_hm_doppler", (8, 32, 48), 2
    monkeypatch.setattr(ops, "S2D_MIN_VOXELS", 0)
    x, poses, tgt = G.make_example(cfg, batch, grid, seed=211)
    sd = O.synth_state_dict(cfg, seed=3)
    outs = {}
    for share in (True, False):
        monkeypatch.setattr(ops, "USE_S2D_SHARE", share)
        eng, params = build_engine(cfg, sd)
        lib.call_counts.clear()
        outs[share], _, _ = run_engine(eng, params, x, tgt)
        n_fold = lib.call_counts.get("rtp_s2d_fold_wgrad", 0)
        # branch 0 feeds 2 (stage 3) + 3 (stage 4) stride-2 fuse convs, branch 1 two more in stage 4 (when its grid is eligible)
        assert n_fold in ((5, 7) if share else (0,)), n_fold
        assert lib.call_counts.get("rtp_gn_apply_s2d", 0) > 0
    a, b = outs[True], outs[False]
    r_hm, r_reg, r_loss, r_grads = oracle_run(x, sd, cfg, tgt, False)
    b_hm, b_reg, b_loss, b_grads = oracle_run(x, sd, cfg, tgt, True)
    for name, r, au in (("hm", r_hm, b_hm), ("reg", r_reg, b_reg)):
        spread = max((au - r).abs().max().item(), 0.02 * r.std().item())
        e = (a[name] - r).abs().max().item()
        print("%s: shared-view path max err %.4g (per-conv path %.4g, autocast %.4g)" % (name, e, (b[name] - r).abs().max().item(), spread))
        assert e <= 2.0 * spread, (name, e, spread)
    assert abs(a["loss"][0].item() - r_loss) <= 1.5e-2 * abs(r_loss)
    rel, cos = grad_report(a["grads"], r_grads)
    rel_b, cos_b = grad_report(b["grads"], r_grads)
    brel, bcos = grad_report(b_grads, r_grads)
    print("grad vs oracle: shared cosine %.5f rel-L2 median %.3g | per-conv %.5f %.3g | autocast %.5f %.3g" %
          (cos, np.median(rel), cos_b, np.median(rel_b), bcos, np.median(brel)))
    assert set(a["grads"]) >= set(r_grads)
    assert cos >= min(0.97, bcos - 0.01) and np.median(rel) <= 1.5 * np.median(brel)
    # the parameters whose gradients take the new route, one by one against the oracle (yardstick: the per-conv path)
    for k in sorted(r_grads):
        if ".fuse_layers." in k and k.split(".fuse_layers.")[1].split(".")[1:3] in (["0", "0"], ["1", "0"]):
            r = r_grads[k].double().flatten()
            ea = float((a["grads"][k].double().flatten() - r).norm() / r.norm())
            eb = float((b["grads"][k].double().flatten() - r).norm() / r.norm())
            print("  %-60s rel-L2 shared %.3g per-conv %.3g" % (k, ea, eb))
            assert ea <= max(2.0 * eb, 0.05), (k, ea, eb)
