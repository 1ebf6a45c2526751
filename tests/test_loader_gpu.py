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


def test_cube_loader_surfaces_reader_errors(tmp_path):
    from rtpose_b200 import lib, loader
    paths, _ = _files(tmp_path, 2, (2, 32, 128, 256), 0, 1, seed=0)
    with open(paths[1], "r+b") as f:
        f.truncate(1000)
    it = iter(loader.CubeLoader(paths, batch=1))
    next(it)
    with pytest.raises(lib.RtpError, match="header promises"):
        next(it)
