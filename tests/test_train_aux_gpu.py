"""GPU parity for the two "next" rows: device-side target assignment (N1) against the CPU oracle
(which is pinned to the reference's AssignLabelPose/2 + gaussian3D), and the fused clip + decoupled-decay + Adam step
(N2) against torch's own clip_grad_norm_ / Adam on CPU driven exactly like OptimWrapper.step."""
import numpy as np
import pytest
import torch

from oracle import hrpose_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("one_hm", [True, False])
def test_device_targets_match_oracle(one_hm):
    from rtpose_b200 import targets
    grid = (16, 64, 160)
    rs = np.random.RandomState(11)
    poses = targets.random_poses(rs, 6, grid)
    poses[2, 0] = [0.78, -5.02, -1.08]   # pelvis in the corner voxel: splat clipped at the border
    poses[4, 5] = [40.0, 0.0, 0.0]       # a joint outside the ROI
    ref = O.batch_targets(list(poses), grid, one_hm)
    got = targets.assign_device(torch.from_numpy(poses).cuda(), grid, one_hm, min_radius=2 if one_hm else 1)
    torch.cuda.synchronize()
    for k in ("ind", "mask", "cat"):
        assert torch.equal(got[k].cpu(), ref[k]), k
    assert torch.equal(got["anno_pose"].cpu(), ref["anno_pose"])
    hm, rhm = got["hm"].cpu(), ref["hm"]
    assert torch.equal(hm > 0, rhm > 0)
    np.testing.assert_allclose(hm.numpy(), rhm.numpy(), rtol=0, atol=1e-7)


def test_fused_adam_matches_reference_optimizer_semantics():
    from rtpose_b200.optim import FlatAdam, one_cycle
    g = torch.Generator().manual_seed(0)
    shapes = [(32, 32, 3, 3, 3), (32,), (128, 192, 1, 1, 1), (45,)]
    n = sum(int(np.prod(s)) for s in shapes)
    flat0 = torch.randn(n, generator=g) * 0.05
    # CPU reference: per-tensor parameters, clip_grad_norm_(35) -> p *= 1 - wd*lr -> Adam(betas=(mom, 0.99)).step()
    ref_params, o = [], 0
    for s in shapes:
        k = int(np.prod(s))
        ref_params.append(flat0[o:o + k].clone().view(s).requires_grad_(True))
        o += k
    opt = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0)
    p = flat0.clone().cuda()
    gr = torch.zeros_like(p)
    fa = FlatAdam(p, gr, wd=0.01, max_norm=35.0)
    for step in range(4):
        lr, mom = one_cycle(step, 10, lr_max=2e-3)
        grads = torch.randn(n, generator=g) * (30.0 if step == 1 else 0.5)   # step 1 exceeds max_norm -> clipping active
        o = 0
        for t in ref_params:
            t.grad = grads[o:o + t.numel()].clone().view(t.shape)
            o += t.numel()
        total = torch.nn.utils.clip_grad_norm_(ref_params, max_norm=35, norm_type=2)
        with torch.no_grad():
            for t in ref_params:
                t.mul_(1 - 0.01 * lr)
        for grp in opt.param_groups:
            grp["lr"], grp["betas"] = lr, (mom, 0.99)
        opt.step()
        gr.copy_(grads.cuda())
        fa.step(lr, mom)
        torch.cuda.synchronize()
        assert abs(float(fa.grad_norm) - float(total)) <= 1e-4 * float(total)
        ref_flat = torch.cat([t.detach().flatten() for t in ref_params])
        torch.testing.assert_close(p.cpu(), ref_flat, rtol=2e-5, atol=2e-7)
