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


def test_fused_adam_follows_reference_trajectory():
    """rtp_adam_step on the device against the parameter trajectory produced by the reference's own OptimWrapper + OneCycle
    + clip_grad_norm_ (oracle/make_optim_golden.py -> tests/golden/optim_golden.npz; 6 steps, step 1 clipped)."""
    import os
    from rtpose_b200.optim import FlatAdam, one_cycle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim_golden.npz"))
    p = torch.from_numpy(g["p0"].astype(np.float32)).cuda()
    gr = torch.zeros_like(p)
    fa = FlatAdam(p, gr, wd=0.01, max_norm=35.0)
    for step in range(6):
        lr, mom = one_cycle(step, 10, lr_max=0.002)
        gr.copy_(torch.from_numpy(g["grad_%d" % step].astype(np.float32)).cuda())
        fa.step(lr, mom)
        torch.cuda.synchronize()
        ref = torch.from_numpy(g["p_%d" % step])
        assert abs(float(fa.grad_norm) - float(g["norm_%d" % step])) <= 1e-4 * float(g["norm_%d" % step])
        torch.testing.assert_close(p.cpu(), ref, rtol=2e-5, atol=2e-7)


def test_adam_by_value_and_device_hyper_agree():
    from rtpose_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(100003, generator=g) * 0.1
    pa, pb = p0.clone().cuda(), p0.clone().cuda()
    ga, gb = torch.zeros_like(pa), torch.zeros_like(pb)
    a, b = FlatAdam(pa, ga), FlatAdam(pb, gb)
    for step in range(3):
        gr = (torch.randn(100003, generator=g) * (40.0 if step == 1 else 0.3)).cuda()
        ga.copy_(gr); gb.copy_(gr)
        a.step(1e-3 * (step + 1), 0.9 - 0.02 * step)            # rtp_adam_step_dev
        b.step_by_value(1e-3 * (step + 1), 0.9 - 0.02 * step)    # rtp_adam_step
    torch.cuda.synchronize()
    torch.testing.assert_close(pa, pb, rtol=1e-6, atol=1e-8)
    assert float(a.grad_norm) == float(b.grad_norm)


def test_step_graph_replay_matches_eager():
    """wgrad on the side stream + join + Adam with device hyper-parameters, captured once and replayed along a
    schedule, equals the same steps issued eagerly (rtpose_b200/graph.py)."""
    from rtpose_b200 import ops
    from rtpose_b200.graph import StepGraph
    from rtpose_b200.optim import FlatAdam
    from rtpose_b200.p8 import P8
    g = torch.Generator().manual_seed(11)
    xp = P8.from_ncdhw(torch.randn(2, 32, 4, 12, 10, generator=g).cuda())
    dyp = P8.from_ncdhw(torch.randn(2, 32, 4, 12, 10, generator=g).cuda())
    sched = [(1e-3, 0.95), (2e-3, 0.9), (5e-4, 0.85)]

    def make():
        p = torch.full((32 * 32 * 27,), 0.01, device="cuda")
        gr = torch.zeros_like(p)
        opt = FlatAdam(p, gr)

        def body():
            ops.conv_wgrad_async(xp, dyp, 3, 1, gr.view(32, 32, 3, 3, 3))
            ops.join_wgrad()
            opt.step_dev()
            return opt.grad_norm
        return p, opt, body

    p_e, opt_e, body_e = make()
    for lr, mom in sched:
        opt_e.set_hyper(lr, mom)
        body_e()
    p_g, opt_g, body_g = make()
    sg = StepGraph(body_g, warmup=0)
    for lr, mom in sched:
        opt_g.set_hyper(lr, mom)
        norm = sg()
    torch.cuda.synchronize()
    assert float(norm) > 0
    assert torch.equal(p_e, p_g)
