"""CenterNet-style training targets for the radar-pose head (host side, numpy).

Mirrors det3d/datasets/pipelines/pose.py:186-255 (`AssignLabelPose`, 15 heatmaps / M=15) and :385-452
(`AssignLabelPose2`, one pelvis heatmap + 45-d offsets / M=1) together with det3d/core/utils/center_utils.py:67-91
(`gaussian3D` with its (2 sigma^2)^(3/2) denominator, `draw_gaussian3D` max-splat).  In the reference this runs in
DataLoader workers; it is the row ranked N1 ("next") in SURVEY.md §8f and stays on the host here.
"""
import numpy as np

VOXEL_SIZE = (0.0453125, 0.15703125, 0.3625)  # x, y, z  (configs/cruw_pose/hr3d.py:101,127-129)
PC_RANGE = (0.7703125, -5.0250000000000234, -1.0875000000000021)  # x, y, z minima of roi1 (hr3d.py:31-33)


def _gaussian(radius):
    d = 2 * radius + 1
    sigma = d / 6
    m = (d - 1.0) / 2.0
    z, y, x = np.ogrid[-m:m + 1, -m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y + z * z) / (2 * sigma * sigma) ** (3 / 2))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def _splat(hm, cx, cy, cz, radius):
    g = _gaussian(radius)
    Z, Y, X = hm.shape
    x0, x1 = min(cx, radius), min(X - cx, radius + 1)
    y0, y1 = min(cy, radius), min(Y - cy, radius + 1)
    z0, z1 = min(cz, radius), min(Z - cz, radius + 1)
    dst = hm[cz - z0:cz + z1, cy - y0:cy + y1, cx - x0:cx + x1]
    src = g[radius - z0:radius + z1, radius - y0:radius + y1, radius - x0:radius + x1]
    if min(src.shape) > 0 and min(dst.shape) > 0:
        np.maximum(dst, src, out=dst)


def assign(poses, grid_zyx, one_hm, min_radius, voxel=VOXEL_SIZE, pc_range=PC_RANGE):
    """poses: [B, 15, 3] metres (one pose per frame, max_poses=1).  Returns dict of numpy arrays
    hm [B,ncls,Z,Y,X] f32, ind [B,M] i64, mask [B,M] u8, cat [B,M] i64, anno_pose [B,M,R] f32."""
    poses = np.asarray(poses, dtype=np.float64)
    B = poses.shape[0]
    Z, Y, X = grid_zyx
    lo = np.array([pc_range[2], pc_range[1], pc_range[0]], dtype=np.float32)  # z, y, x (float32 as in the reference)
    ncls, M, R = (1, 1, 45) if one_hm else (15, 15, 3)
    out = dict(hm=np.zeros((B, ncls, Z, Y, X), np.float32), ind=np.zeros((B, M), np.int64),
               mask=np.zeros((B, M), np.uint8), cat=np.zeros((B, M), np.int64), anno_pose=np.zeros((B, M, R), np.float32))
    for b in range(B):
        ct = np.empty((15, 3), dtype=np.float32)
        for i in range(15):
            x, y, z = poses[b, i]
            ct[i] = ((x - lo[2]) / voxel[0], (y - lo[1]) / voxel[1], (z - lo[0]) / voxel[2])
        ci = ct.astype(np.int32)
        if one_hm:
            cx, cy, cz = ci[0]
            if 0 <= cx < X and 0 <= cy < Y and 0 <= cz < Z:
                _splat(out["hm"][b, 0], cx, cy, cz, min_radius)
                out["ind"][b, 0] = cz * Y * X + cy * X + cx
                out["mask"][b, 0] = 1
                out["anno_pose"][b, 0] = (ct - ci[0][None, :].astype(np.float32)).flatten()
        else:
            for k in range(15):
                cx, cy, cz = ci[k]
                if not (0 <= cx < X and 0 <= cy < Y and 0 <= cz < Z):
                    continue
                _splat(out["hm"][b, k], cx, cy, cz, max(min_radius, 1))
                out["cat"][b, k] = k
                out["ind"][b, k] = cz * Y * X + cy * X + cx
                out["mask"][b, k] = 1
                out["anno_pose"][b, k] = ct[k] - ci[k].astype(np.float32)
    return out


def random_poses(rs, batch, grid_zyx, voxel=VOXEL_SIZE, pc_range=PC_RANGE):
    """Random 15-joint skeletons: pelvis uniform in the ROI shrunk by 0.5 m, joints = pelvis + N(0, 0.3 m)."""
    Z, Y, X = grid_zyx
    lo = np.array(pc_range, dtype=np.float64)
    ext = np.array([X * voxel[0], Y * voxel[1], Z * voxel[2]])
    shrink = np.minimum(0.5, 0.25 * ext)
    out = np.empty((batch, 15, 3))
    for b in range(batch):
        pelvis = lo + shrink + rs.uniform(0, 1, 3) * (ext - 2 * shrink)
        j = pelvis[None] + rs.normal(0, 0.3, (15, 3))
        j[0] = pelvis
        out[b] = np.clip(j, lo + 1e-3, lo + ext - 1e-3)
    return out


def assign_device(poses, grid_zyx, one_hm, min_radius, voxel=VOXEL_SIZE, pc_range=PC_RANGE):
    """Same targets, built on the GPU by rtp_assign_targets: poses is a CUDA float64 tensor
    [B, 15, 3]; returns CUDA tensors hm, ind, mask, cat, anno_pose with the reference's dtypes and shapes."""
    import ctypes as C

    import torch

    from . import lib
    assert poses.is_cuda and poses.dtype == torch.float64 and poses.shape[1:] == (15, 3)
    poses = poses.contiguous()
    B = poses.shape[0]
    Z, Y, X = grid_zyx
    ncls, M, R = (1, 1, 45) if one_hm else (15, 15, 3)
    dev = poses.device
    out = dict(hm=torch.empty((B, ncls, Z, Y, X), dtype=torch.float32, device=dev),
               ind=torch.empty((B, M), dtype=torch.int64, device=dev), mask=torch.empty((B, M), dtype=torch.uint8, device=dev),
               cat=torch.empty((B, M), dtype=torch.int64, device=dev), anno_pose=torch.empty((B, M, R), dtype=torch.float32, device=dev))
    v = (C.c_double * 3)(*[float(a) for a in voxel])
    r = (C.c_float * 3)(*[float(a) for a in pc_range])
    radius = min_radius if one_hm else max(min_radius, 1)
    lib.call("rtp_assign_targets", poses.data_ptr(), B, Z, Y, X, 1 if one_hm else 0, int(radius), v, r, out["hm"].data_ptr(),
             out["ind"].data_ptr(), out["mask"].data_ptr(), out["cat"].data_ptr(), out["anno_pose"].data_ptr(),
             torch.cuda.current_stream().cuda_stream)
    return out
