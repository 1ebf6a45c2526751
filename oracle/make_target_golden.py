"""TEST INFRASTRUCTURE — tests/golden/targets_golden.npz from the reference's own target assigners
(AssignLabelPose / AssignLabelPose2.__call__, det3d/datasets/pipelines/pose.py:153-255, :345-452, with gaussian3D /
draw_gaussian3D of det3d/core/utils/center_utils.py:67-91), run UNMODIFIED from their source files behind import shims.

    python -m oracle.make_target_golden

Promotion rules.  The assigners compute `(x - radar_range[2]) / voxel_size[0] / out_size_factor[2]` with `x` a python float and
`radar_range` a float32 array.  Under the NumPy the reference runs with (1.x: torch==2.0.1 cannot load NumPy 2) a python float
and a float32 SCALAR promote to float64, so the voxel coordinate is computed in double from the float32-rounded range; under
NumPy >= 2 (NEP 50, this container has 2.3) the same expression stays in float32 and `anno_pose` moves by up to 2 float32 ulps
(1.5e-5; heat-maps, indices, masks and categories are identical — measured on 400 random poses).  NumPy 2 has no switch for
the legacy rules, so the module's `np.array` is wrapped to hand that ONE array (the only 2-D float32 `np.array` call in the
file) back as float64 holding the float32-rounded values, which reproduces the legacy arithmetic exactly; everything else runs
as written.
"""
import importlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "targets_golden.npz")
NAMES = ["Pelvis", "Right_Hip", "Right_Knee", "Right_Ankle", "Left_Hip", "Left_Knee", "Left_Ankle", "Thomx", "Head",
         "Left_Shoulder", "Left_Elbow", "Left_Wrist", "Right_Shoulder", "Right_Elbow", "Right_Wrist"]
ROI = {"z": [-1.0875000000000021, 4.7125], "y": [-5.0250000000000234, 5.024999999999931], "x": [0.7703125, 8.0203125]}  # configs/cruw_pose/hr3d.py:31
GRID_SIZE = [0.0453125, 0.15703125, 0.3625]


class A(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class _LegacyNp:
    """numpy, except that a 2-D float32 np.array(...) comes back as float64 of the float32-rounded values (see module doc)."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def array(obj, *a, **kw):
        r = np.array(obj, *a, **kw)
        return r.astype(np.float64) if (r.ndim == 2 and r.dtype == np.float32) else r


def load_pose_module(legacy=True):
    def mod(name, path=None, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = [path] if path else []
        m.__package__ = name
        sys.modules[name] = m
        return m

    class _Reg:
        def register_module(self, cls):
            return cls

    mod("det3d"); mod("det3d.core"); mod("det3d.core.bbox", box_np_ops=None); mod("det3d.core.sampler", preprocess=None)
    mod("det3d.builder", build_dbsampler=None); mod("det3d.core.input"); mod("det3d.core.input.voxel_generator", VoxelGenerator=None)
    mod("det3d.core.utils", REF + "/det3d/core/utils"); mod("det3d.core.utils.circle_nms_jit", circle_nms=None)
    mod("det3d.datasets"); mod("det3d.datasets.registry", PIPELINES=_Reg())
    mod("det3d.datasets.pipelines", REF + "/det3d/datasets/pipelines")
    pose = importlib.import_module("det3d.datasets.pipelines.pose")
    pose.np = _LegacyNp() if legacy else np
    return pose


def run_reference(pose_mod, poses, grid, one_hm):
    ncls = 1 if one_hm else 15
    cfg = A(out_size_factor=[1, 1, 1], target_assigner=A(tasks=[A(num_class=ncls, class_names=NAMES[:ncls])]), gaussian_overlap=0.1,
            max_poses=1, min_radius=2 if one_hm else 1)  # configs/cruw_pose/hr3d_one_hm_doppler.py:104-105, hr3d.py:107-108
    info = A(DATASET=A(ROI=A(roi1=ROI), LABEL=A(ROI_TYPE="roi1"), RDR_CUBE=A(GRID_SIZE=GRID_SIZE)))
    cls = pose_mod.AssignLabelPose2 if one_hm else pose_mod.AssignLabelPose
    res = {"rdr_cube": np.zeros((1,) + tuple(grid), np.float32), "mode": "train", "hm_size": tuple(grid),
           "poses": [np.asarray(p).tolist() for p in poses], "meta": {}}
    out, _ = cls(cfg=cfg)(res, info)
    return {k: np.asarray(out["rdr"][k][0]) for k in ("hm", "ind", "mask", "cat", "anno_pose")}


def synth_poses(n, grid, seed=0):
    """Random skeletons inside the ROI, plus one whose pelvis lies outside it (the assigner must skip it)."""
    from oracle import hrpose_oracle as O
    rs = np.random.RandomState(seed)
    poses = [O.synth_pose(rs, grid) for _ in range(n - 1)]
    out = poses[0].copy()
    out[:, 0] += 20.0
    return poses + [out]


def main():
    grid = (16, 64, 160)
    pose_mod = load_pose_module(legacy=True)
    poses = synth_poses(8, grid)
    pack = {"poses": np.stack(poses), "grid": np.array(grid)}
    for one_hm in (True, False):
        tag = "one_hm" if one_hm else "hr3d"
        for i, p in enumerate(poses):
            r = run_reference(pose_mod, [p], grid, one_hm)
            nz = np.flatnonzero(r["hm"])
            pack["%s_%d_hm_idx" % (tag, i)], pack["%s_%d_hm_val" % (tag, i)] = nz.astype(np.int64), r["hm"].reshape(-1)[nz]
            pack["%s_%d_hm_shape" % (tag, i)] = np.array(r["hm"].shape)
            for k in ("ind", "mask", "cat", "anno_pose"):
                pack["%s_%d_%s" % (tag, i, k)] = r[k]
    np.savez_compressed(OUT, **pack)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
