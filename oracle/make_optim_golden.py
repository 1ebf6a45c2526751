"""TEST INFRASTRUCTURE — tests/golden/optim_golden.npz: parameter trajectories produced by the reference's own optimizer stack,
run from its source files: `OptimWrapper` (det3d/solver/fastai_optim.py:121-260) created the way `build_one_cycle_optimizer`
does (det3d/torchie/apis/train.py:148-174: Adam(betas=(0.9, 0.99)), wd, true_wd=fixed_wd, bn_wd=True, one layer group of the
flattened model), driven by `OneCycle` (det3d/solver/learning_schedules_fastai.py:77-95) and `OptimizerHook.clip_grads`
(det3d/torchie/trainer/hooks/optimizer.py:9-12: clip_grad_norm_(max_norm=35, norm_type=2)) with the cruw_pose config values
(configs/cruw_pose/hr3d_one_hm_doppler.py:170-179).  Gradients are seeded random tensors; step 1 is scaled to trip the clip.

    python -m oracle.make_optim_golden

Only environment shim: `collections.Iterable` (removed in Python 3.10; fastai_optim.py:1 imports it).
"""
import collections
import collections.abc
import importlib.util
import os
from functools import partial

import numpy as np
import torch
from torch import nn

REF = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "optim_golden.npz")
STEPS, TOTAL = 6, 10


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def flatten_model(m):  # det3d/torchie/apis/train.py:148-153
    return sum(map(flatten_model, m.children()), []) if len(list(m.children())) else [m]


def make_model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv3d(4, 8, 3, bias=False), nn.GroupNorm(4, 8), nn.ReLU(), nn.Conv3d(8, 4, 1, bias=True))


def main():
    collections.Iterable = collections.abc.Iterable
    fo = _load("ref_fastai_optim", "det3d/solver/fastai_optim.py")
    ls = _load("ref_sched", "det3d/solver/learning_schedules_fastai.py")
    model = make_model()
    params = [p for p in model.parameters()]
    opt = fo.OptimWrapper.create(partial(torch.optim.Adam, betas=(0.9, 0.99), amsgrad=0.0), 3e-3, [nn.Sequential(*flatten_model(model))],
                                 wd=0.01, true_wd=True, bn_wd=True)
    sched = ls.OneCycle(opt, TOTAL, 0.002, [0.95, 0.85], 10.0, 0.4)
    g = torch.Generator().manual_seed(1)
    pack = {"p0": torch.cat([p.detach().flatten() for p in params]).numpy().copy(),
            "shapes": np.array([list(p.shape) + [0] * (5 - p.dim()) for p in params]), "ndims": np.array([p.dim() for p in params])}
    for step in range(STEPS):
        sched.step(step)  # LrUpdaterHook: before the iteration
        grads = [torch.randn(p.shape, generator=g) * (60.0 if step == 1 else 0.5) for p in params]
        opt.zero_grad()
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        total = torch.nn.utils.clip_grad.clip_grad_norm_(filter(lambda p: p.requires_grad, model.parameters()), max_norm=35, norm_type=2)
        opt.step()
        pack["grad_%d" % step] = torch.cat([gr.flatten() for gr in grads]).numpy()
        pack["norm_%d" % step] = np.array(float(total))
        pack["lr_mom_%d" % step] = np.array([opt.lr, opt.mom])
        pack["p_%d" % step] = torch.cat([p.detach().flatten() for p in params]).numpy().copy()
    np.savez_compressed(OUT, **pack)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; norms", [round(float(pack["norm_%d" % s]), 2) for s in range(STEPS)])


if __name__ == "__main__":
    main()
