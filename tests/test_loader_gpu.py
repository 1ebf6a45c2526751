"""GPU parity of the file -> model-input path: CubeLoader (ROI-row reads -> pinned slab -> device ->
rtp_ingest_pack) against the oracle's restatement of CRUW_POSE_Dataset.get_cube / get_cube_phase applied to np.load of
the same files.  The fp32 side output must be bit-identical to numpy; the P8 tensor is its bf16 rounding."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16).float()


def _files(tmp_path, n, shape, lo, hi, seed):
    rs = np.random.RandomState(seed)
    paths, arrs = [], []
    for i in range(n):
        a = rs.uniform(lo, hi, size=shape).astype(np.float16)
        p = str(tmp_path / ("%06d.npy" % i))
        np.save(p, a)
        paths.append(p)
        arrs.append(a)
    return paths, arrs


@pytest.mark.parametrize("kind", ["doppler", "zyx", "phase"])
def test_cube_loader_matches_reference_ingest(tmp_path, kind):
    from oracle import hrpose_oracle as O
    from rtpose_b200 import loader
    if kind == "doppler":      # dzyx_real: [D,32,128,256], normalise by (0, 10)  (configs/cruw_pose/hr3d_one_hm_doppler.py)
        shape, norm, lo, hi = (8, 32, 128, 256), (0.0, 10.0), -2, 12
    elif kind == "zyx":        # zyx_real: [32,128,256] -> one channel; (a, b) scaled into the finite fp16 range
        shape, norm, lo, hi = (32, 128, 256), (30000.0, 50000.0), 25000, 60000
    else:                      # complex cube [2,D,32,128,256] -> 2D channels, crop only (get_cube_phase)
        shape, norm, lo, hi = (2, 4, 32, 128, 256), None, -1, 1
    paths, arrs = _files(tmp_path, 7, shape, lo, hi, seed=len(kind))
    ld = loader.CubeLoader(paths, batch=3, norm=norm, depth=2, frame_workers=2, io_threads=3, drop_last=False, want_f32=True)
    assert len(ld) == 3
    seen = 0
    for (x, f32), ps in ld:
        n = len(ps)
        assert ps == paths[seen:seen + n]
        if kind == "phase":
            ref = np.stack([O.ingest_cube_phase(a) for a in arrs[seen:seen + n]])
        elif kind == "zyx":
            ref = np.stack([O.ingest_cube(a, norm) for a in arrs[seen:seen + n]])  # ingest_cube adds the channel axis
        else:
            ref = np.stack([O.ingest_cube(a, norm) for a in arrs[seen:seen + n]])
        torch.cuda.synchronize()
        assert (x.N, x.C) == (n, ref.shape[1])
        assert np.array_equal(f32.cpu().numpy(), ref), "fp32 side output must be bit-identical to numpy"
        assert torch.equal(x.to_ncdhw().cpu(), bf(torch.from_numpy(ref)))
        seen += n
    assert seen == 7
    # a second epoch over the same loader gives the same first batch (fresh staging, no state carried over)
    (x2, f2), _ = next(iter(ld))
    torch.cuda.synchronize()
    assert np.array_equal(f2.cpu().numpy()[0], O.ingest_cube_phase(arrs[0]) if kind == "phase" else O.ingest_cube(arrs[0], norm))
    ld.close()  # the second epoch was abandoned after one batch


def test_cube_loader_surfaces_reader_errors(tmp_path):
    from rtpose_b200 import lib, loader
    paths, _ = _files(tmp_path, 2, (2, 32, 128, 256), 0, 1, seed=0)
    with open(paths[1], "r+b") as f:
        f.truncate(1000)
    it = iter(loader.CubeLoader(paths, batch=1))
    next(it)
    with pytest.raises(lib.RtpError, match="header promises"):
        next(it)
    with pytest.raises(StopIteration):
        next(it)


def test_detector_accepts_the_loaders_packed_input(tmp_path):
    """RadarPoseNet.forward(example) with example['rdr']['rdr_tensor'] = the P8 the loader yields == the same call with
    the reference's fp32 tensor (bf16 rounding happens at the same place either way), eager and graphed, train and test."""
    from oracle import hrpose_oracle as O
    from rtpose_b200 import det3d_compat as D
    from rtpose_b200 import loader, targets
    B, in_ch, grid = 2, 32, (16, 64, 160)
    paths, arrs = _files(tmp_path, B, (in_ch, 32, 128, 256), -2, 12, seed=4)
    names = ["Pelvis"]
    cfg = dict(type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
               backbone=dict(type="HRNet3D", backbone_cfg="hr_tiny_feat32_zyx_l4_in32", final_conv_in=192, final_conv_out=128,
                             final_fuse="conat_conv", ds_factor=1),
               pose_head=dict(type="CenterHead", tasks=[dict(num_class=1, class_names=names)], in_channels=128,
                              share_conv_channel=128, dataset="cruw_pose", weight=1.0, code_weights=[1.0] * 45,
                              common_heads={"reg": (45, 2)}, dcn_head=False), neck=None)
    torch.manual_seed(0)
    class TestCfg(dict):
        __getattr__ = dict.__getitem__
    test_cfg = TestCfg(post_center_limit_range=[], score_threshold=0.0, pc_range=list(O.PC_RANGE), out_size_factor=[1, 1, 1],
                       voxel_size=list(O.VOXEL_SIZE))
    model = D.build_detector(cfg, train_cfg=None, test_cfg=test_cfg).cuda()
    rs = np.random.RandomState(2)
    tg = targets.assign(targets.random_poses(rs, B, grid), grid, one_hm=True, min_radius=2)
    (xp, f32), _ = next(iter(loader.CubeLoader(paths, batch=B, norm=(0.0, 10.0), want_f32=True)))

    def example(x):
        ex = {"rdr": {"rdr_tensor": x}, "meta": [{}] * B}
        for k, v in tg.items():
            ex["rdr"][k] = [torch.from_numpy(v).cuda()]
        return ex

    def step(x, graph):
        model.cuda_graph = graph
        for p in model.parameters():
            p.grad = None
        losses = model(example(x), return_loss=True)
        losses["loss"][0].backward()
        torch.cuda.synchronize()
        return float(losses["loss"][0]), torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()

    l_ref, g_ref = step(f32, False)
    for graph in (False, True, True):
        l, g = step(xp, graph)
        assert l == l_ref and torch.equal(g, g_ref), "packed input must give the same bits (graph=%s)" % graph
    model.eval()
    with torch.no_grad():
        a = model(example(f32), return_loss=False)
        b = model(example(xp), return_loss=False)
    assert a[0]["keypoints"] == b[0]["keypoints"] and a[1]["keypoints"] == b[1]["keypoints"]
