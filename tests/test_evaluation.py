"""Evaluation post-step.  CPU: the oracle's restatement against the golden produced by the unmodified
reference code (oracle/make_eval_golden.py -> tests/golden/eval_golden.json).  GPU: rtp_pjpe / rtp_pjpe_seq_mean and the
`evaluation` drop-in against the same golden and against the oracle on random inputs (fp64; per-frame errors bit-exact)."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_golden.json")


def _flat(res):
    out = dict(("ALL/" + k, v) for k, v in res["results"].items())
    for s, d in res["seq_results"].items():
        out.update((s + "/" + k, v) for k, v in d.items())
    return out


def test_oracle_evaluation_equals_reference_golden():
    from oracle import hrpose_oracle as O
    g = json.load(open(GOLD))
    mine, ref = _flat(O.evaluation(g["detections"], g["gt"], g["seq_id_to_name"])), _flat(g["result"])
    assert mine.keys() == ref.keys() and len(ref) == 4 * 32
    for k in ref:
        assert mine[k] == ref[k], k  # same numpy operations in the same order: identical doubles


@pytest.mark.gpu
def test_device_evaluation_equals_reference_golden():
    from rtpose_b200 import evaluation as E
    g = json.load(open(GOLD))
    res, extra = E.evaluation(g["detections"], g["gt"], g["seq_id_to_name"])
    assert extra is None
    mine, ref = _flat(res), _flat(g["result"])
    assert mine.keys() == ref.keys()
    for k in ref:
        assert abs(mine[k] - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (k, mine[k], ref[k])


@pytest.mark.gpu
def test_device_pjpe_is_bit_exact_per_frame():
    from oracle import hrpose_oracle as O
    from rtpose_b200 import evaluation as E
    rs = np.random.RandomState(0)
    N = 257
    gt = rs.uniform(-5, 8, size=(N, 15, 3))
    pred = (gt + rs.normal(0, 0.1, size=gt.shape)).astype(np.float32)
    pred[3] = gt[3].astype(np.float32)                     # (near-)zero errors
    rel, ab = E.pjpe(torch.from_numpy(pred).cuda().view(N, 1, 45), torch.from_numpy(gt).cuda())  # one_hm row layout
    want_rel = np.stack([O.pjpe(pred[n], gt[n]) for n in range(N)])
    want_abs = np.stack([O.abs_pjpe(pred[n], gt[n]) for n in range(N)])
    assert np.array_equal(rel.cpu().numpy(), want_rel) and np.array_equal(ab.cpu().numpy(), want_abs)
    assert np.all(rel.cpu().numpy()[:, 0] == 0.0)          # the root joint is the origin of the relative metric
    seq = rs.randint(0, 4, size=N)
    seq[seq == 2] = 3                                       # sequence 2 has no frames
    r_mm, a_mm, cnt = E.sequence_means(rel, ab, torch.from_numpy(seq), 4)
    for s in range(4):
        sel = seq == s
        assert int(cnt[s]) == sel.sum()
        if sel.any():
            np.testing.assert_allclose(r_mm[s].cpu().numpy(), want_rel[sel].mean(axis=0) * 1000, rtol=1e-13)
            np.testing.assert_allclose(a_mm[s].cpu().numpy(), want_abs[sel].mean(axis=0) * 1000, rtol=1e-13)
        else:
            assert float(r_mm[s].abs().sum()) == 0.0


@pytest.mark.gpu
def test_evaluation_rejects_cpu_tensors():
    from rtpose_b200 import evaluation as E
    from rtpose_b200 import lib
    with pytest.raises(lib.RtpError, match="CUDA"):
        E.pjpe(torch.zeros(1, 15, 3), torch.zeros(1, 15, 3, dtype=torch.float64))
