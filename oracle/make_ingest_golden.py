"""TEST INFRASTRUCTURE — tests/golden/ingest_golden.npz from the reference dataset's own cube path, run from source:
`consider_roi_cube` / `get_arr_in_roi` (det3d/datasets/cruw_pose/cruw_pose.py:125-146) on the axis vectors of
`__init__` (:38-40) and the config ROI (configs/cruw_pose/hr3d.py:31), then `get_cube` (:167-185) and `get_cube_phase`
(:188-194) called unbound on a stub carrying the attributes they read.  The hard-coded `/mnt/ssd3/cruw_pose/...` np.load is
served by a wrapper that returns a seeded synthetic float16 cube (the dataset itself is not available).

    python -m oracle.make_ingest_golden

Stored: the ROI index list, and for each cube kind a strided sub-sample of the result plus its float64 sum / abs-sum / count of
zeros (the clamp); tests regenerate the same input from the seed.
"""
import os
import types

import numpy as np

from oracle.make_eval_golden import load_dataset_class

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ingest_golden.npz")
ROI = {"z": [-1.0875000000000021, 4.7125], "y": [-5.0250000000000234, 5.024999999999931], "x": [0.7703125, 8.0203125]}
SUB = (slice(None, None, 3), slice(None, None, 5), slice(None, None, 7))  # z, y, x sub-sampling of the stored result


def synth_cube(kind, seed):
    """float16 cube as stored on disk: 'dzyx' [D,32,128,256] in ~[-2, 12] (normalised by (0, 10)), 'zyx' [32,128,256] in a
    range that survives float16, 'phase' [2,D,32,128,256] in [-1, 1]."""
    rs = np.random.RandomState(seed)
    if kind == "dzyx":
        return rs.uniform(-2, 12, size=(3, 32, 128, 256)).astype(np.float16)
    if kind == "zyx":
        return rs.uniform(25000, 60000, size=(32, 128, 256)).astype(np.float16)
    return rs.uniform(-1, 1, size=(2, 2, 32, 128, 256)).astype(np.float16)


class _NpWithLoad:
    def __init__(self, arr):
        self._arr = arr

    def __getattr__(self, k):
        return getattr(np, k)

    def load(self, path, *a, **kw):
        assert path.startswith("/mnt/ssd3/cruw_pose") and path.endswith(".npy"), path
        return self._arr.copy()


def main():
    cls = load_dataset_class()
    glob = cls.get_cube.__globals__  # the dataset module's namespace: its `np` is swapped while get_cube runs
    stub = types.SimpleNamespace(arr_z_cb=np.arange(-5.8, 5.8, 11.6 / 32), arr_y_cb=np.arange(-10.05, 10.05, 20.1 / 128),
                                 arr_x_cb=np.arange(0, 11.6, 11.6 / 256), seq_id_to_name={"0": "2024_0218_1209"})
    stub.get_arr_in_roi = types.MethodType(cls.get_arr_in_roi, stub)
    cls.consider_roi_cube(stub, ROI)
    pack = {"roi_idx": np.array(stub.list_roi_idx_cb, dtype=np.int64)}
    A = lambda **kw: types.SimpleNamespace(**kw)
    for kind, rdr_type, norm, seed in (("dzyx", "dzyx_real", ["0", "10"], 21), ("zyx", "zyx_real", [30000.0, 50000.0], 22),
                                       ("phase", "dzyx_complex", None, 23)):
        raw = synth_cube(kind, seed)
        glob["np"] = _NpWithLoad(raw)
        stub.cfg = A(DATASET=A(RDR_TYPE=rdr_type))
        stub.rad_normalize_values = norm
        out = cls.get_cube_phase(stub, "0", "000123") if kind == "phase" else cls.get_cube(stub, "0", "000123")
        glob["np"] = np
        out = np.asarray(out)
        flat = out.reshape((-1,) + out.shape[-3:])
        pack[kind + "_shape"], pack[kind + "_dtype"] = np.array(out.shape), np.array(str(out.dtype))
        pack[kind + "_sub"] = flat[(slice(None),) + SUB].copy()
        pack[kind + "_stats"] = np.array([flat.astype(np.float64).sum(), np.abs(flat.astype(np.float64)).sum(), float((flat == 0).sum())])
        pack[kind + "_seed"] = np.array(seed)
    np.savez_compressed(OUT, **pack)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; roi", pack["roi_idx"].tolist(), {k: pack[k + "_shape"].tolist() for k in ("dzyx", "zyx", "phase")})


if __name__ == "__main__":
    main()
