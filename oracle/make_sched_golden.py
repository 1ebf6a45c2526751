"""TEST INFRASTRUCTURE — tests/golden/one_cycle_golden.json from the reference's own OneCycle scheduler
(det3d/solver/learning_schedules_fastai.py:53-95), loaded from its source file (pure numpy, no det3d imports).

    python -m oracle.make_sched_golden
"""
import importlib.util
import json
import os
import types

REF = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "one_cycle_golden.json")


def main():
    spec = importlib.util.spec_from_file_location("ref_sched", os.path.join(REF, "det3d/solver/learning_schedules_fastai.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    cases = []
    # (total_step, lr_max, moms, div_factor, pct_start): the cruw_pose configs' values and two odd splits
    for total, lr_max, moms, div, pct in ((1000, 0.002, [0.95, 0.85], 10.0, 0.4), (37, 0.01, [0.9, 0.8], 25.0, 0.3),
                                          (10, 0.002, [0.95, 0.85], 10.0, 0.4)):
        opt = types.SimpleNamespace(lr=None, mom=None)
        sched = m.OneCycle(opt, total, lr_max, moms, div, pct)
        vals = []
        for step in range(total):
            sched.step(step)
            vals.append([float(opt.lr), float(opt.mom)])
        cases.append({"total_step": total, "lr_max": lr_max, "moms": moms, "div_factor": div, "pct_start": pct, "lr_mom": vals})
    json.dump(cases, open(OUT, "w"))
    print("wrote", OUT, [len(c["lr_mom"]) for c in cases])


if __name__ == "__main__":
    main()
