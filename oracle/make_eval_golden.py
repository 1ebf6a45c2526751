"""TEST INFRASTRUCTURE — generates tests/golden/eval_golden.json by running the UNMODIFIED reference evaluation code
(CRUW_POSE_Dataset.evaluation, det3d/datasets/cruw_pose/cruw_pose.py:277-310, with PJPE / ABS_PJPE of eval_util.py)
on synthetic detections.  Runs only in the build container (needs /root/reference); the fixture travels.

    python -m oracle.make_eval_golden

The dataset module is loaded behind shims for what is not installed (munch) or not needed (dataset registry, pipelines);
`evaluation` is called unbound on a stub that carries the two attributes it reads (label_file, seq_id_to_name).
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eval_golden.json")


def load_dataset_class():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    class _Reg:
        def register_module(self, cls):
            return cls

    mod("munch", DefaultMunch=type("DefaultMunch", (), {"fromDict": staticmethod(lambda d: d)}))
    for n in ("det3d", "det3d.datasets"):
        if n not in sys.modules:
            mod(n)
    mod("det3d.datasets.registry", DATASETS=_Reg())
    mod("det3d.datasets.pipelines", Compose=object)
    sys.path.insert(0, REF)  # for `from eval_util import *`
    spec = importlib.util.spec_from_file_location("ref_cruw_pose", os.path.join(REF, "det3d/datasets/cruw_pose/cruw_pose.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CRUW_POSE_Dataset


def synth(seed=0):
    """3 sequences with 5 / 3 / 1 frames; predictions are fp32 values (as .cpu() tensors give), labels are doubles."""
    rs = np.random.RandomState(seed)
    names = {"0": "2024_0218_1209", "1": "2024_0301_0930", "7": "2024_0302_1411"}
    gt, det = {}, {}
    for seq, nframes in (("0", 5), ("1", 3), ("7", 1)):
        gt[seq] = {}
        for f in range(nframes):
            pose = rs.uniform([0.8, -5.0, -1.0], [8.0, 5.0, 4.7], size=(15, 3))
            pred = (pose + rs.normal(0, 0.05, size=(15, 3))).astype(np.float32)
            frame = "%06d" % (f * 3)
            gt[seq][frame] = [{"pose": pose.tolist()}]
            det["%s/%s/%06d" % (seq, frame, f * 3 + 1)] = {
                "keypoints": [(j, float(pred[j, 0]), float(pred[j, 1]), float(pred[j, 2]), float(rs.rand())) for j in range(15)],
                "metadata": {}}
    return names, gt, det


def main():
    cls = load_dataset_class()
    names, gt, det = synth()
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(gt, f)
        label_file = f.name
    stub = types.SimpleNamespace(label_file=label_file, seq_id_to_name=names)
    res, _ = cls.evaluation(stub, {k: {"keypoints": [tuple(p) for p in v["keypoints"]], "metadata": {}} for k, v in det.items()})
    os.unlink(label_file)
    to_py = lambda d: {k: (to_py(v) if isinstance(v, dict) else float(v)) for k, v in d.items()}
    json.dump({"seq_id_to_name": names, "gt": gt, "detections": det, "result": to_py(res)}, open(OUT, "w"))
    print("wrote", OUT, "MPJPE", res["results"]["MPJPE"], "ABS_MPJPE", res["results"]["ABS_MPJPE"])


if __name__ == "__main__":
    main()
