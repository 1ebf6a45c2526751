"""On-disk cube reader: rtp_npy_probe / rtp_npy_read_roi_slab are host code in the C-ABI library, so
they are checked here without a GPU against numpy's own reader and the reference's crop
(CRUW_POSE_Dataset.get_cube / get_cube_phase, det3d/datasets/cruw_pose/cruw_pose.py:170-192)."""
import os
import threading

import numpy as np
import pytest
import torch

from rtpose_b200 import lib, loader


def _save(path, arr, version=None):
    if version is None:
        np.save(path, arr)
    else:
        with open(path, "wb") as f:
            np.lib.format.write_array(f, arr, version=version)
    return str(path)


@pytest.mark.parametrize("shape,roi", [((6, 10, 12, 24), (3, 5, 2, 7)),      # Doppler cube [D,RZ,RY,RX]
                                       ((10, 12, 24), (0, 10, 0, 12)),       # 3-D cube, ROI = everything (merged runs)
                                       ((2, 3, 10, 12, 24), (9, 1, 11, 1)),  # complex cube [2,D,RZ,RY,RX], last row
                                       ((5, 32, 128, 256), (13, 16, 32, 64))])  # the real geometry, 5 Doppler bins
@pytest.mark.parametrize("threads", [1, 5])
def test_roi_slab_equals_numpy_crop(tmp_path, shape, roi, threads):
    rs = np.random.RandomState(len(shape) + threads)
    arr = rs.uniform(-3, 3, size=shape).astype(np.float16)
    p = _save(tmp_path / "cube.npy", arr)
    info = loader.probe(p)
    assert info["shape"] == shape and info["descr"] == "<f2" and not info["fortran_order"]
    assert info["file_bytes"] == os.path.getsize(p) and info["data_offset"] + arr.nbytes == info["file_bytes"]
    z0, Z, y0, Y = roi
    got = loader.read_roi_slab(p, z0, Z, y0, Y, threads=threads).numpy()
    want = arr.reshape((-1,) + shape[-3:])[:, z0:z0 + Z, y0:y0 + Y, :]
    assert got.shape == want.shape and np.array_equal(got.view(np.uint16), want.view(np.uint16))  # bit-exact, NaN-safe


def test_slab_then_x_crop_is_the_reference_crop(tmp_path):
    """slab[..., x0:x0+X] == arr[:, z0:z1+1, y0:y1+1, x0:x1+1] of get_cube (the x crop happens in rtp_ingest_pack)."""
    from oracle import hrpose_oracle as O
    arr = np.random.RandomState(0).uniform(0, 12, size=(3, 32, 128, 256)).astype(np.float16)
    p = _save(tmp_path / "f.npy", arr)
    z0, z1, y0, y1, x0, x1 = O.ROI_IDX
    slab = loader.read_roi_slab(p).numpy()  # defaults = the reference ROI
    assert np.array_equal(slab[..., x0:x1 + 1], arr[:, z0:z1 + 1, y0:y1 + 1, x0:x1 + 1])
    ref = O.ingest_cube(arr, (0.0, 10.0))
    mine = (slab[..., x0:x1 + 1].astype(np.float32) - 0.0) / 10.0
    mine[mine < 0] = 0
    assert np.array_equal(mine, ref)


@pytest.mark.parametrize("version", [(1, 0), (2, 0), (3, 0)])
def test_npy_header_versions(tmp_path, version):
    arr = np.arange(4 * 6 * 8, dtype=np.float16).reshape(4, 6, 8)
    p = _save(tmp_path / "v.npy", arr, version)
    assert loader.probe(p)["shape"] == (4, 6, 8)
    assert np.array_equal(loader.read_roi_slab(p, 1, 2, 3, 3).numpy()[0], arr[1:3, 3:6, :])


def test_into_pinned_style_buffer_and_concurrent_calls(tmp_path):
    arrs = [np.random.RandomState(i).uniform(-1, 1, size=(4, 8, 16, 32)).astype(np.float16) for i in range(6)]
    paths = [_save(tmp_path / ("%d.npy" % i), a) for i, a in enumerate(arrs)]
    stage = torch.zeros((6, 4, 4, 8, 32), dtype=torch.float16)
    ts = [threading.Thread(target=loader.read_roi_slab, args=(paths[i], 2, 4, 8, 8, stage[i], 2)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(6):
        assert np.array_equal(stage[i].numpy(), arrs[i][:, 2:6, 8:16, :])


def test_reader_errors(tmp_path):
    good = np.zeros((2, 4, 6, 8), dtype=np.float16)
    p = _save(tmp_path / "g.npy", good)
    with pytest.raises(lib.RtpError, match="cannot open"):
        loader.probe(str(tmp_path / "missing.npy"))
    with pytest.raises(lib.RtpError, match="outside the cube"):
        loader.read_roi_slab(p, 2, 3, 0, 6)
    with pytest.raises(lib.RtpError, match="outside the cube"):
        loader.read_roi_slab(p, 0, 4, -1, 2)
    with pytest.raises(lib.RtpError, match="destination holds"):
        loader.read_roi_slab(p, 0, 4, 0, 6, out=torch.empty(10, dtype=torch.float16))
    with pytest.raises(lib.RtpError, match="contiguous CPU float16"):
        loader.read_roi_slab(p, 0, 4, 0, 6, out=torch.empty((2, 4, 6, 8), dtype=torch.float32))
    p32 = _save(tmp_path / "f32.npy", np.zeros((2, 4, 6, 8), dtype=np.float32))
    with pytest.raises(lib.RtpError, match="float16"):
        loader.read_roi_slab(p32, 0, 4, 0, 6, out=torch.empty(2 * 4 * 6 * 8, dtype=torch.float16))
    pf = _save(tmp_path / "fo.npy", np.asfortranarray(np.zeros((4, 6, 8), dtype=np.float16)))
    with pytest.raises(lib.RtpError, match="Fortran"):
        loader.read_roi_slab(pf, 0, 4, 0, 6)
    p2 = _save(tmp_path / "2d.npy", np.zeros((6, 8), dtype=np.float16))
    with pytest.raises(lib.RtpError, match="at least"):
        loader.read_roi_slab(p2, 0, 1, 0, 1, out=torch.empty(64, dtype=torch.float16))
    junk = tmp_path / "junk.npy"
    junk.write_bytes(b"not a numpy file at all")
    with pytest.raises(lib.RtpError, match="bad magic"):
        loader.probe(str(junk))
    data = open(p, "rb").read()
    cut = tmp_path / "cut.npy"
    cut.write_bytes(data[:-100])
    with pytest.raises(lib.RtpError, match="header promises"):
        loader.probe(str(cut))
    empty = tmp_path / "empty.npy"
    empty.write_bytes(b"")
    with pytest.raises(lib.RtpError, match="shorter than"):
        loader.probe(str(empty))


def test_loader_refuses_to_run_without_a_device(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    p = _save(tmp_path / "g.npy", np.zeros((2, 32, 128, 256), dtype=np.float16))
    with pytest.raises(lib.RtpError, match="no CPU fallback"):
        loader.CubeLoader([p], 1)


def test_shard_paths_gives_every_rank_the_same_number_of_whole_batches():
    paths = ["f%03d.npy" % i for i in range(103)]
    for world, batch in ((1, 16), (2, 16), (4, 8), (8, 3), (3, 1)):
        shards = [loader.shard_paths(paths, r, world, batch) for r in range(world)]
        n = (103 // world // batch) * batch
        assert all(len(s) == n for s in shards), (world, batch, [len(s) for s in shards])
        flat = [p for s in shards for p in s]
        assert len(set(flat)) == len(flat) and all(s == sorted(s) for s in shards)   # disjoint, order kept
        for r, s in enumerate(shards):
            if s:
                from rtpose_b200.dist import shard_frames
                assert paths.index(s[0]) == shard_frames(103, r, world)[0]
