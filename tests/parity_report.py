"""Prints the path-level parity numbers: our CUDA path vs the fp32 CPU oracle, next to the oracle's own
bf16-autocast spread on the same inputs/weights (SURVEY.md §8c contract (3)).  Run on the GPU box:

    python tests/parity_report.py [cfg ...]  > gpurun_out/parity.txt

Lives under tests/ because it drives the oracle (test infrastructure); it is a report, not a collected test.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # this directory

from oracle import hrpose_oracle as O  # noqa: E402
from oracle import make_golden as G  # noqa: E402


def oracle_run(x, sd, cfg, tgt, autocast):
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    c = O.CONFIGS[cfg]
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        preds = O.forward(torch.from_numpy(x), sdr, cfg)
    preds = {k: v.float() for k, v in preds.items()}
    L = O.head_loss(preds, tgt, c["weight"], c["code_weights"])
    L["loss"].backward()
    grads = {k: v.grad for k, v in sdr.items() if v.grad is not None}
    return preds, float(L["loss"]), grads


def compare(name, hm, reg, loss, grads, ref):
    from test_engine_gpu import grad_report
    rhm, rreg, rloss, rgrads = ref
    rel, cos = grad_report(grads, {k: v for k, v in rgrads.items() if float(v.norm()) > 0})
    row = dict(who=name, hm_max_err=float((hm - rhm).abs().max()), reg_max_err=float((reg - rreg).abs().max()),
               loss=loss, loss_rel=abs(loss - rloss) / abs(rloss), grad_cos=cos, grad_rel_median=float(np.median(rel)),
               grad_rel_max=float(rel.max()))
    print(json.dumps(row))
    return row


def main():
    from test_engine_gpu import build_engine, run_engine
    cfgs = sys.argv[1:] or ["hr3d", "hr3d_one_hm_doppler", "hr3d_one_hm_doppler_phase"]
    torch.set_num_threads(os.cpu_count() or 8)
    for cfg in cfgs:
        grid, batch = (8, 16, 24), 2
        x, poses, tgt = G.make_example(cfg, batch, grid, seed=101)
        sd = O.synth_state_dict(cfg, seed=3)
        p32, l32, g32 = oracle_run(x, sd, cfg, tgt, False)
        ref = (p32["hm"].detach(), p32["reg"].detach(), l32, g32)
        print("== %s grid %s batch %d: fp32 oracle loss %.5f, hm std %.4f, reg std %.4f" %
              (cfg, grid, batch, l32, float(ref[0].std()), float(ref[1].std())))
        pb, lb, gb = oracle_run(x, sd, cfg, tgt, True)
        compare("oracle bf16-autocast (cpu)", pb["hm"].detach(), pb["reg"].detach(), lb, gb, ref)
        eng, params = build_engine(cfg, sd)
        out, _, _ = run_engine(eng, params, x, tgt)
        compare("rtpose_b200 (cuda)", out["hm"], out["reg"], float(out["loss"][0]), out["grads"], ref)


if __name__ == "__main__":
    main()
