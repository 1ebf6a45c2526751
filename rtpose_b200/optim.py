"""Fused optimizer step on flat buffers.

Mirrors what the reference does per iteration around the hot path:
  OptimizerHook.after_train_iter  det3d/torchie/trainer/hooks/optimizer.py:14-24   clip_grad_norm_(max_norm=35, L2)
  OptimWrapper.step               det3d/solver/fastai_optim.py:158-174             p *= 1 - wd*lr  (true_wd, bn_wd=True)
  torch.optim.Adam(betas=(mom, 0.99)) built at det3d/torchie/apis/train.py:157-174
  OneCycle lr / momentum schedule det3d/solver/learning_schedules_fastai.py:53-95   (values passed in per step)
as ONE pass over (param, grad, m, v) in rtp_adam_step.  Parameters and gradients must live in flat fp32 buffers
(the engine already produces gradients that way).
"""
import math

import numpy as np
import torch

from . import lib
from .p8 import _stream


class FlatAdam:
    """step(lr, mom) = set_hyper(lr, mom) + step_dev().  The kernel reads the schedule point from an 8-float DEVICE block
    (rtp_adam_step_dev), so step_dev() can sit inside a captured CUDA graph while set_hyper() — a 32-byte async copy
    from a pinned ring — runs before each replay."""
    RING = 1024  # pinned schedule slots: the host may run this many steps ahead of the device

    def __init__(self, flat_params, flat_grads, wd=0.01, eps=1e-8, beta2=0.99, max_norm=35.0, params=()):
        """params: tensors that alias the flat buffer WITHOUT sharing its version counter (nn.Parameters whose .data was
        pointed into it); views taken with flat[a:b].view(...) share the counter and need not be listed."""
        assert flat_params.is_cuda and flat_params.dtype == torch.float32 and flat_params.is_contiguous()
        assert flat_grads.shape == flat_params.shape and flat_grads.dtype == torch.float32
        self.p, self.g = flat_params, flat_grads
        self.m, self.v = torch.zeros_like(flat_params), torch.zeros_like(flat_params)
        self.wd, self.eps, self.beta2, self.max_norm = float(wd), float(eps), float(beta2), float(max_norm)
        self.t = 0
        self._aliases = list(params)
        dev = flat_params.device
        self.ws = torch.empty(lib.load().rtp_adam_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._ring = torch.zeros((self.RING, 8), dtype=torch.float32).pin_memory()

    def set_hyper(self, lr, mom=0.9):
        """Advances the Adam time step and uploads {lr, beta1, beta2, eps, wd, 1-beta1^t, sqrt(1-beta2^t), max_norm}."""
        self.t += 1
        slot = self._ring[self.t % self.RING]
        # fp32 like rtp_adam_step: bias1 = 1 - powf(beta1, t), bias2_sqrt = sqrtf(1 - powf(beta2, t))
        f32 = np.float32
        slot[0], slot[1], slot[2], slot[3], slot[4] = float(lr), float(mom), self.beta2, self.eps, self.wd
        slot[5] = float(f32(1) - np.power(f32(mom), f32(self.t)))
        slot[6] = float(np.sqrt(f32(1) - np.power(f32(self.beta2), f32(self.t))))
        slot[7] = self.max_norm
        self.hyper.copy_(slot, non_blocking=True)

    def _mark_written(self):
        """The kernel rewrote the parameters through raw pointers: tell torch, so that everything keyed on a parameter's
        `_version` — the engine's bf16 weight packs (ops.PackedWeights), its space-to-depth and merged-head keys — sees
        the change and repacks.  (Inside a replayed CUDA graph no Python runs: a captured step must contain
        `packs.refresh_async()` itself, as bench.py and det3d_compat do.)"""
        torch.autograd.graph.increment_version(self.p)
        for t in self._aliases:
            torch.autograd.graph.increment_version(t)

    def step_dev(self):
        lib.call("rtp_adam_step_dev", self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.p.numel(),
                 self.hyper.data_ptr(), self.ws.data_ptr(), self.grad_norm.data_ptr(), _stream())
        self._mark_written()

    def step(self, lr, mom=0.9):
        self.set_hyper(lr, mom)
        self.step_dev()

    def step_by_value(self, lr, mom=0.9):
        """Same update through rtp_adam_step (hyper-parameters as kernel arguments)."""
        self.t += 1
        lib.call("rtp_adam_step", self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.p.numel(),
                 float(lr), float(mom), self.beta2, self.eps, self.wd, self.t, self.max_norm, self.ws.data_ptr(),
                 self.grad_norm.data_ptr(), _stream())
        self._mark_written()


def one_cycle(step, total_steps, lr_max=2e-3, div_factor=10.0, pct_start=0.4, moms=(0.95, 0.85)):
    """OneCycle (learning_schedules_fastai.py:53-95): cosine warm-up lr_max/div -> lr_max over pct_start of the run with
    momentum moms[0] -> moms[1], then cosine annealing to lr_max/div/1e4 and back to moms[0]."""
    a1 = int(total_steps * pct_start)
    a2 = total_steps - a1
    low = lr_max / div_factor

    def cos(start, end, pct):
        return end + (start - end) / 2 * (math.cos(math.pi * pct) + 1)

    if step < a1:
        pct = step / max(a1, 1)
        return cos(low, lr_max, pct), cos(moms[0], moms[1], pct)
    pct = (step - a1) / max(a2, 1)
    return cos(lr_max, low / 1e4, pct), cos(moms[1], moms[0], pct)
