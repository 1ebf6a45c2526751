"""Pose-error evaluation with the per-frame work on the device.

Mirrors `CRUW_POSE_Dataset.evaluation` (det3d/datasets/cruw_pose/cruw_pose.py:277-310) and `PJPE` / `ABS_PJPE`
(eval_util.py:5-11): same inputs (the `detections` dict `CenterHead.predict` produces, keyed 'seq/frame/rdr_frame', and
the label file's dict), same result structure `(res, None)` with `res['results']` / `res['seq_results']`.
`pjpe` / `sequence_means` are the tensor-level entry points for a test loop that keeps the decoded joints on the device
(rtp_decode's `out_xyz`).  Kernels: rtp_pjpe, rtp_pjpe_seq_mean (fp64, the reference's operation order).  The last step
— means over 15 joints and over sequences, a few hundred doubles — is numpy on the host, as in the reference.
"""
import json

import numpy as np
import torch

from . import lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pjpe(pred_xyz, gt_xyz):
    """pred_xyz CUDA fp32 [N,J,3] (or [N,1,3J]); gt_xyz CUDA fp64 [N,J,3] -> (root-relative, absolute) errors, fp64 [N,J], metres."""
    lib.require_device()
    if not (torch.is_tensor(pred_xyz) and pred_xyz.is_cuda and torch.is_tensor(gt_xyz) and gt_xyz.is_cuda):
        raise lib.RtpError("pjpe: inputs must be CUDA tensors (no CPU fallback)")
    N, J = gt_xyz.shape[0], gt_xyz.shape[1]
    pred = pred_xyz.to(torch.float32).contiguous()
    gt = gt_xyz.to(torch.float64).contiguous()
    if pred.numel() != N * J * 3 or gt.shape[2] != 3:
        raise lib.RtpError("pjpe: pred has %d values, gt is %r" % (pred.numel(), tuple(gt.shape)))
    rel = torch.empty((N, J), dtype=torch.float64, device=gt.device)
    ab = torch.empty_like(rel)
    lib.call("rtp_pjpe", pred.data_ptr(), gt.data_ptr(), N, J, rel.data_ptr(), ab.data_ptr(), _stream())
    return rel, ab


def sequence_means(rel, ab, seq_index, num_seq):
    """Per-sequence, per-joint means x 1000 (mm): fp64 [S,J] twice, and the frame count per sequence (int32 [S])."""
    N, J = rel.shape
    seq_index = seq_index.to(device=rel.device, dtype=torch.int32).contiguous()
    if seq_index.numel() != N:
        raise lib.RtpError("sequence_means: %d sequence indices for %d frames" % (seq_index.numel(), N))
    rel_mm = torch.empty((num_seq, J), dtype=torch.float64, device=rel.device)
    ab_mm = torch.empty_like(rel_mm)
    count = torch.empty((num_seq,), dtype=torch.int32, device=rel.device)
    lib.call("rtp_pjpe_seq_mean", rel.data_ptr(), ab.data_ptr(), seq_index.data_ptr(), N, J, num_seq, rel_mm.data_ptr(),
             ab_mm.data_ptr(), count.data_ptr(), _stream())
    return rel_mm, ab_mm, count


def summarize(rel_mm, ab_mm, seq_names):
    """[S,J] per-sequence means (host numpy) -> the reference's result dict (cruw_pose.py:290-310)."""
    seq_res = {}
    for s, name in enumerate(seq_names):
        r, a = rel_mm[s], ab_mm[s]
        out = {"MPJPE": np.mean(r), "ABS_MPJPE": np.mean(a)}
        for j in range(r.shape[0]):
            out["PJPE_%d" % j], out["ABS_PJPE_%d" % j] = r[j], a[j]
        seq_res[name] = out
    total = {"MPJPE": np.mean([v["MPJPE"] for v in seq_res.values()]), "ABS_MPJPE": np.mean([v["ABS_MPJPE"] for v in seq_res.values()])}
    for i in range(15):
        total["PJPE_%d" % i] = np.mean([v["PJPE_%d" % i] for v in seq_res.values()])
        total["ABS_PJPE_%d" % i] = np.mean([v["ABS_PJPE_%d" % i] for v in seq_res.values()])
    seq_res["ALL"] = total
    return {"results": total, "seq_results": seq_res}


def evaluation(detections, gt, seq_id_to_name, device="cuda", output_dir=None, testset=False):
    """Drop-in for `dataset.evaluation(detections)`: `gt` is the label dict or the path of the label file."""
    lib.require_device()
    if isinstance(gt, str):
        with open(gt, "r") as f:
            gt = json.load(f)
    seqs, seq_idx, pred, lab = [], [], [], []
    for key, val in detections.items():
        seq, frame, _ = key.split("/")
        if seq not in seqs:
            seqs.append(seq)  # first-appearance order, as the reference's defaultdict
        seq_idx.append(seqs.index(seq))
        pred.append([p[1:4] for p in val["keypoints"]])
        lab.append(gt[seq][frame][0]["pose"])
    if not pred:
        raise lib.RtpError("evaluation: no detections")
    # predictions are fp32 values that went through python floats: the fp32 cast is exact
    pred_t = torch.tensor(pred, dtype=torch.float64).to(torch.float32).to(device)
    gt_t = torch.tensor(lab, dtype=torch.float64, device=device)
    rel, ab = pjpe(pred_t, gt_t)
    rel_mm, ab_mm, _ = sequence_means(rel, ab, torch.tensor(seq_idx, dtype=torch.int32), len(seqs))
    return summarize(rel_mm.cpu().numpy(), ab_mm.cpu().numpy(), [seq_id_to_name[s] for s in seqs]), None
