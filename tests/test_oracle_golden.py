"""The CPU oracle (oracle/hrpose_oracle.py) against golden vectors produced by the reference's own modules
(oracle/make_golden.py, run in the build container).  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O
from oracle import make_golden as G

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_g*x*x*.npz")))  # model goldens (make_golden.py)


def _load(path):
    g = np.load(path, allow_pickle=False)
    cfg, batch, grid, seed = [str(v) for v in g["meta"]]
    return g, cfg, int(batch), tuple(int(v) for v in grid.split("x")), int(seed)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_forward_loss_grads_decode_match_reference(path):
    g, cfg, batch, grid, seed = _load(path)
    torch.set_num_threads(8)
    x, poses, tgt = G.make_example(cfg, batch, grid, seed)
    assert abs(float(x.astype(np.float64).sum()) - float(g["x_checksum"])) < 1e-6
    np.testing.assert_array_equal(poses, g["poses"])
    for k in ("hm", "ind", "mask", "cat", "anno_pose"):
        np.testing.assert_array_equal(tgt[k].numpy(), g["tgt_" + k])
    sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(cfg).items()}
    preds = O.forward(torch.from_numpy(x), sd, cfg)
    # same ATen kernels on both sides: only summation-order noise is allowed
    np.testing.assert_allclose(preds["hm"].detach().numpy(), g["hm"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(preds["reg"].detach().numpy(), g["reg"], rtol=1e-4, atol=1e-4)
    c = O.CONFIGS[cfg]
    L = O.head_loss(preds, tgt, c["weight"], c["code_weights"])
    assert abs(float(L["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(float(L["hm_loss"]) - float(g["hm_loss"])) <= 1e-5 * abs(float(g["hm_loss"]))
    np.testing.assert_allclose(L["loc_loss_elem"].detach().numpy(), g["loc_loss_elem"], rtol=1e-4, atol=1e-6)
    assert float(L["num_positive"]) == float(g["num_positive"])
    L["loss"].backward()
    names = [str(n) for n in g["grad_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    for k, ref in norms.items():
        got = float(sd[k].grad.norm()) if sd[k].grad is not None else 0.0
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-6) + 1e-7, (k, got, ref)
    for k in g.files:
        if k.startswith("grad::"):
            np.testing.assert_allclose(sd[k[6:]].grad.numpy(), g[k], rtol=2e-3, atol=2e-5 * float(np.abs(g[k]).max() + 1e-12) + 1e-9)
    # decode: integer indices exact, coordinates to fp32 round-off
    kps, idxs = O.decode(preds["hm"].detach(), preds["reg"].detach())
    for n in range(batch):
        assert len(kps[n]) == int(g["num_keypoints"][n])
        ref = g["keypoints"][n][: len(kps[n])]
        got = np.array(kps[n], dtype=np.float64)
        np.testing.assert_array_equal(got[:, 0], ref[:, 0])
        np.testing.assert_allclose(got[:, 1:], ref[:, 1:], rtol=1e-5, atol=1e-5)


def test_state_dict_spec_counts():
    # parameter counts measured on the reference (SURVEY.md §6 / BASELINE.md §2)
    want = {"hr3d": 2002194, "hr3d_one_hm": 2217006, "hr3d_one_hm_doppler": 2216942,
            "hr3d_one_hm_doppler_phase": 8297518}
    for cfg, n in want.items():
        got = sum(int(np.prod(s)) for _, s in O.state_dict_spec(cfg))
        assert got == n, (cfg, got, n)


def test_ingest_matches_numpy_semantics():
    rs = np.random.RandomState(0)
    raw = (rs.uniform(-2, 12, size=(4, 32, 128, 256))).astype(np.float16)
    out = O.ingest_cube(raw, (0.0, 10.0))
    assert out.shape == (4, 16, 64, 160) and out.dtype == np.float32
    ref = (raw.astype(np.float32)[:, 13:29, 32:96, 17:177] - 0.0) / 10.0
    ref[ref < 0] = 0
    np.testing.assert_array_equal(out, ref)
    one = O.ingest_cube(raw[0], (150000.0, 200000.0))
    assert one.shape == (1, 16, 64, 160)


def test_target_assignment_equals_reference_golden():
    """oracle.assign_targets against the reference's own AssignLabelPose / AssignLabelPose2 (pose.py:153-255, :345-452) run
    from source by oracle/make_target_golden.py (NumPy-1 promotion, see that file): heat-map, ind, mask, cat and anno_pose
    bit-exact for 7 random skeletons and one outside the ROI, both label layouts."""
    from oracle import hrpose_oracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "targets_golden.npz"))
    grid = tuple(int(v) for v in g["grid"])
    assert grid == (16, 64, 160) and g["poses"].shape == (8, 15, 3)
    for tag, one_hm in (("one_hm", True), ("hr3d", False)):
        for i, pose in enumerate(g["poses"]):
            mine = O.assign_targets(pose, grid, one_hm, radius=2 if one_hm else 1)
            hm = np.zeros(tuple(g["%s_%d_hm_shape" % (tag, i)]), np.float32)
            hm.reshape(-1)[g["%s_%d_hm_idx" % (tag, i)]] = g["%s_%d_hm_val" % (tag, i)]
            assert hm.shape == mine["hm"].shape and np.array_equal(hm, mine["hm"]), (tag, i, "hm")
            for k in ("ind", "mask", "cat", "anno_pose"):
                ref = g["%s_%d_%s" % (tag, i, k)]
                assert ref.dtype == np.asarray(mine[k]).dtype and np.array_equal(ref, mine[k]), (tag, i, k)
    assert int(g["one_hm_7_mask"].sum()) == 0 and int(g["hr3d_7_mask"].sum()) < 15  # the skeleton pushed out of the ROI


def test_fused_optimizer_step_equals_reference_trajectory():
    """oracle.adam_step_flat (the arithmetic of rtp_adam_step) driven by rtpose_b200.optim.one_cycle against the parameter
    trajectory the reference's own OptimWrapper + OneCycle + clip_grad_norm_ produced (oracle/make_optim_golden.py ->
    tests/golden/optim_golden.npz): 6 steps, the second one clipped (norm 1789 > 35).  float32 both sides; the operation
    order differs slightly (torch divides by bias corrections separately), hence rel 2e-6 on the parameters."""
    from rtpose_b200.optim import one_cycle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim_golden.npz"))
    p = g["p0"].astype(np.float32).copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for step in range(6):
        lr, mom = one_cycle(step, 10, lr_max=0.002)
        assert abs(lr - g["lr_mom_%d" % step][0]) <= 1e-12 and abs(mom - g["lr_mom_%d" % step][1]) <= 1e-12
        norm = O.adam_step_flat(p, g["grad_%d" % step].astype(np.float32), m, v, step + 1, lr, mom)
        assert abs(norm - float(g["norm_%d" % step])) <= 1e-5 * float(g["norm_%d" % step])
        ref = g["p_%d" % step]
        assert np.abs(p - ref).max() <= 2e-6 * np.abs(ref).max() + 1e-8, (step, np.abs(p - ref).max())
    assert float(g["norm_1"]) > 35.0 > float(g["norm_0"])


def test_ingest_equals_reference_dataset_golden():
    """oracle.ROI_IDX / ingest_cube / ingest_cube_phase against the reference dataset's own consider_roi_cube + get_cube +
    get_cube_phase run from source (oracle/make_ingest_golden.py -> tests/golden/ingest_golden.npz), bit-exact on a strided
    sub-sample plus sum / abs-sum / zero-count of the whole result."""
    from oracle import make_ingest_golden as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz"))
    assert tuple(int(v) for v in g["roi_idx"]) == tuple(O.ROI_IDX) == (13, 28, 32, 95, 17, 176)
    for kind, norm in (("dzyx", (0.0, 10.0)), ("zyx", (30000.0, 50000.0)), ("phase", None)):
        raw = M.synth_cube(kind, int(g[kind + "_seed"]))
        mine = O.ingest_cube_phase(raw) if kind == "phase" else O.ingest_cube(raw, norm)
        assert str(g[kind + "_dtype"]) == "float32" and mine.dtype == np.float32
        assert int(np.prod(g[kind + "_shape"])) == mine.size and tuple(g[kind + "_shape"][-3:]) == mine.shape[-3:] == (16, 64, 160)
        flat = mine.reshape((-1,) + mine.shape[-3:])
        assert np.array_equal(flat[(slice(None),) + M.SUB], g[kind + "_sub"]), kind
        stats = [flat.astype(np.float64).sum(), np.abs(flat.astype(np.float64)).sum(), float((flat == 0).sum())]
        assert stats == g[kind + "_stats"].tolist(), (kind, stats, g[kind + "_stats"].tolist())
    assert g["dzyx_stats"][2] > 0  # the clamp was exercised


def test_space_to_depth_identity_of_stride2_convs():
    """DESIGN.md §3.5 on the CPU in float64: conv3d(x, W, stride 2, pad 1) == conv3d(s2d_view(x), s2d_expand_weight(W),
    stride 1, pad 1), and 19 of the 27 x 8 (tap, parity) blocks of the expanded weight are structurally zero."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, 6, 10, generator=g, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.conv3d(x, w, stride=2, padding=1)
    we = O.s2d_expand_weight(w)
    got = torch.nn.functional.conv3d(O.s2d_view(x), we, stride=1, padding=1)
    assert got.shape == ref.shape == (2, 7, 4, 3, 5)
    torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-12)
    blocks = we.reshape(7, 8, 5, 27).abs().sum(dim=(0, 2)) > 0   # [parity, view tap]
    assert int(blocks.sum()) == 27 and blocks.shape == (8, 27)
    # the gradient of the expanded weight folds back onto the 27 original taps (weight_s2d_fold_kernel)
    we_g = we.clone().requires_grad_(True)
    torch.nn.functional.conv3d(O.s2d_view(x), we_g, stride=1, padding=1).square().sum().backward()
    w_g = w.clone().requires_grad_(True)
    torch.nn.functional.conv3d(x, w_g, stride=2, padding=1).square().sum().backward()
    pt = {0: (1, 0), 1: (0, 1), 2: (1, 1)}
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                (pz, tz), (py, ty), (px, tx) = pt[kz], pt[ky], pt[kx]
                par = pz * 4 + px * 2 + py
                torch.testing.assert_close(we_g.grad[:, par * 5:(par + 1) * 5, tz, ty, tx], w_g.grad[:, :, kz, ky, kx], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("cfg", ["hr3d_one_hm_doppler", "hr3d", "hr3d_one_hm_doppler_phase"])
def test_reference_init_distribution_matches_reference_model(cfg):
    """oracle.reference_init_state_dict draws from the distributions the reference's constructors use: compared, tensor by
    tensor, with the state_dict of the reference's own model built under manual_seed(0) (names, order, shapes; constants
    exactly; random tensors by standard deviation and range)."""
    from oracle import make_golden as G
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not available")
    mods = ref_loader.load()
    torch.manual_seed(0)
    model = mods["build_detector"](G.ref_model_cfg(cfg), train_cfg=None, test_cfg=G.ref_test_cfg())
    ref, mine = model.state_dict(), O.reference_init_state_dict(cfg)
    assert list(ref) == list(mine)
    for k in ref:
        a, b = ref[k].float(), mine[k]
        assert a.shape == b.shape, k
        if a.numel() == 1 or float(a.std()) == 0.0:
            assert torch.equal(a, b), k
            continue
        tol = 0.05 if a.numel() >= 8192 else (0.12 if a.numel() >= 512 else 0.6)
        assert abs(float(a.std()) - float(b.std())) <= tol * float(a.std()), (k, float(a.std()), float(b.std()))
        assert abs(float(a.abs().max()) - float(b.abs().max())) <= max(tol, 0.25) * float(a.abs().max()), k
