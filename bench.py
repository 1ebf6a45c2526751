#!/usr/bin/env python
"""bench.py — radar frames/s of the HRRadarPose forward+backward hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cfg hr3d_one_hm_doppler]

A "step" = one pass of the hot path over one batch of synthetic radar cubes: raw fp16 cube + 3-D skeletons (resident in
HBM) -> ingest (ROI crop / normalise / clamp / channel pack) + target assignment -> HRNet3D backbone -> CenterHead ->
focal + L1 loss -> backward (all parameter gradients) [-> NCCL all-reduce of the flat gradient buffer when N > 1] ->
fused clip + Adam step (weights really change, so the bf16 weight packs are rebuilt every step).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAW_SHAPE = (32, 128, 256)  # z, y, x of the on-disk cube (cruw_pose.py:38-40)
ROI0 = (13, 32, 17)         # z0, y0, x0 of roi1 -> 16 x 64 x 160
GRID = (16, 64, 160)
CFGS = {
    # name: (arch, final_in, final_out, fuse, reg, ncls, weight, in_ch, norm(a, b), fwd GFLOP/frame)
    "hr3d": ("hr_tiny_feat32_zyx_l4", 32, 32, "top", 3, 15, 0.2, 1, (150000.0, 200000.0), 114.92),
    "hr3d_one_hm": ("hr_tiny_feat32_zyx_l4", 192, 128, "conat_conv", 45, 1, 0.5, 1, (150000.0, 200000.0), 185.25),
    "hr3d_one_hm_doppler": ("hr_tiny_feat32_zyx_l4_in32", 192, 128, "conat_conv", 45, 1, 0.5, 32, (0.0, 10.0), 185.24),
    "hr3d_one_hm_doppler_phase": ("hr_tiny_feat64_zyx_l4_in64", 384, 256, "conat_conv", 45, 1, 0.7, 64, None, 556.95),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples the SM clock and the throttle reasons while the timed region runs: NVML every 10 ms (a step is ~25 ms,
    the timed region a few hundred ms), falling back to polling nvidia-smi when pynvml is unavailable."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.max_mhz, self.source = index, [], False, None, None

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = int(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        bits = [N.nvmlClocksThrottleReasonHwSlowdown, N.nvmlClocksThrottleReasonHwThermalSlowdown,
                N.nvmlClocksThrottleReasonSwThermalSlowdown, N.nvmlClocksThrottleReasonSwPowerCap]
        self.source = "nvml"
        while not self.stop_flag:
            mhz = int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
            mask = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.rows.append((mhz, [bool(mask & b) for b in bits]))
            time.sleep(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                c = [v.strip() for v in r.stdout.strip().split(",")] if r.returncode == 0 else []
                if len(c) >= 6 and c[0].isdigit():
                    self.max_mhz = int(c[1]) if c[1].isdigit() else self.max_mhz
                    self.rows.append((int(c[0]), [v.lower().startswith("active") for v in c[2:6]]))
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            if not self.stop_flag:
                self._run_smi()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[1][i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.rows),
                "source": self.source}


def workload_config(cfg, B, world, D):
    """The `config` keys that name the workload; shared by both arms so the driver compares like with like."""
    return {"workload": "%s training fwd+bwd, batch %d per GPU, raw fp16 cube [%d,32,128,256] -> ingest -> "
                        "HRNet3D -> CenterHead -> loss -> backward" % (cfg, B, D),
            "per_gpu_batch": B, "global_batch": B * world, "grid": list(GRID), "parallelism": "dp%d" % world}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_runner(cfg, batch, device="cpu", seed=1234):
    """(train_step() -> loss, infer_step() -> detections, kind).  kind "reference": the reference's own, unmodified
    RadarPoseNet (det3d registry -> build_detector) loaded from the staged copy baseline/_ref (or /root/reference in the
    build container) through oracle/ref_loader.py, its stock code path (model(example, return_loss=True) -> loss dict ->
    backward; model(example, return_loss=False) -> predict), random-init weights under manual_seed(0).
    kind "port": the oracle restatement (same ATen ops), used only when no reference tree is reachable."""
    from oracle import hrpose_oracle as O
    from oracle import make_golden as G
    from oracle import ref_loader
    x, poses, tgt = G.make_example(cfg, batch, GRID, seed=seed)
    xt = torch.from_numpy(x).to(device)
    if ref_loader.available():
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):  # the reference prints while importing / building; stdout carries the JSON line
            mods = ref_loader.load()
            torch.manual_seed(0)
            model = mods["build_detector"](G.ref_model_cfg(cfg), train_cfg=None, test_cfg=G.ref_test_cfg()).to(device)
        example = {"rdr": {"rdr_tensor": xt, "hm": [tgt["hm"].to(device)], "anno_pose": [tgt["anno_pose"].to(device)],
                           "ind": [tgt["ind"].to(device)], "mask": [tgt["mask"].to(device)], "cat": [tgt["cat"].to(device)]},
                   "meta": [{"i": i} for i in range(batch)]}

        def train_step():
            model.train()
            model.zero_grad(set_to_none=True)
            losses = model(example, return_loss=True)
            losses["loss"][0].backward()
            return losses["loss"][0]

        def infer_step():
            model.eval()
            with torch.no_grad():
                return model(example, return_loss=False)
        return train_step, infer_step, "reference"
    sd = {k: v.to(device).requires_grad_(True) for k, v in O.synth_state_dict(cfg).items()}

    def train_step():
        for v in sd.values():
            v.grad = None
        L = O.forward_loss(xt, sd, cfg, tgt)
        L["loss"].backward()
        return L["loss"]

    def infer_step():
        with torch.no_grad():
            hm, reg = O.forward(xt, sd, cfg)
            return O.decode(hm, reg)
    return train_step, infer_step, "port"


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (its unmodified modules, fp32, all host threads) on a bounded
    sample: batch 1 forward+backward per step at the full 16x64x160 grid."""
    if rank != 0:
        return
    cfg = args.cfg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, _, kind = reference_runner(cfg, 1)
    warm = max(1, args.warmup)
    for _ in range(warm):
        step()
    steps = max(1, args.steps)  # exactly the K requested: a step is ~0.6 s on 16 cores, K = 10 ends in seconds
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = 1.0 / dt
    line = {"impl": "reference", "metric": "radar frames/sec HRRadarPose fwd+bwd", "value": val, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(cfg, args.batch, max(1, args.gpus), CFGS[cfg][7]),
                           reference_sample="each step = batch 1 forward+backward of that workload on the host cores"),
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "%d steps x batch 1, fp32, %s" % (steps, "the reference's RadarPoseNet from baseline/_ref (stock code path)"
                                                                        if kind == "reference" else "oracle port (same ATen kernels the reference calls)")},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def ref_gpu_probe(args):
    """Informative baseline: the UNMODIFIED reference modules on the same B200 through stock
    PyTorch (cuDNN / ATen), fp32 and bf16 autocast — training fwd+bwd at the bench batch and inference + decode at batch
    32.  Prints one JSON line; not the target, not part of the bench contract."""
    dev = torch.device("cuda", 0)
    out = {"probe": "reference modules on stock PyTorch %s, cuDNN %s, %s" % (torch.__version__, torch.backends.cudnn.version(),
                                                                              torch.cuda.get_device_name(0)), "cfg": args.cfg}
    torch.backends.cudnn.benchmark = True

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for name, B, train in (("train_b%d" % args.batch, args.batch, True), ("infer_b32_incl_decode", 32, False)):
        try:
            tstep, istep, kind = reference_runner(args.cfg, B, device=dev)
            out["kind"] = kind
            fn = tstep if train else istep
            for mode in ("fp32", "tf32", "bf16_autocast"):
                torch.backends.cudnn.allow_tf32 = mode != "fp32"
                torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
                if mode == "bf16_autocast":
                    def run(fn=fn):
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return fn()
                else:
                    run = fn
                ms = timed(run, 5)
                out["%s_%s" % (name, mode)] = {"ms_per_step": round(ms, 2), "frames_per_s": round(B / ms * 1e3, 1)}
            out["%s_peak_mem_GB" % name] = round(torch.cuda.max_memory_allocated() / 1e9, 1)
            del tstep, istep, fn
            torch.cuda.empty_cache()
        except Exception as ex:
            out[name] = {"error": repr(ex)[:300]}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm
def build_params(cfg, device):
    """Random-init parameters of the named architecture in ONE flat fp32 buffer (+ matching flat grad buffer)."""
    from rtpose_b200 import spec
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, gf = CFGS[cfg]
    entries = [("backbone." + n, s, i) for n, s, i in spec.backbone_spec(arch, fin, fout)]
    heads = {"reg": (reg, 2), "hm": (ncls, 2)}
    entries += [("pose_head." + n, s, i) for n, s, i in spec.head_spec(fout if fuse != "top" else 32, fout if fuse != "top" else 32, heads)]
    # reg.0 and hm.0 are evaluated as one merged conv: place their weights (and biases) next to each other in the flat buffer,
    # so the merged operand is a view of it (engine._adjacent_cat) instead of a per-step torch.cat
    order = {"pose_head.tasks.0.reg.0.weight": 0, "pose_head.tasks.0.hm.0.weight": 1, "pose_head.tasks.0.reg.0.bias": 2,
             "pose_head.tasks.0.hm.0.bias": 3}
    tail = sorted((e for e in entries if e[0] in order), key=lambda e: order[e[0]])
    entries = [e for e in entries if e[0] not in order] + tail
    total = sum(int(np.prod(s)) for _, s, _ in entries)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    gflat = torch.zeros(total, dtype=torch.float32, device=device)
    params, grads, o = {}, {}, 0
    torch.manual_seed(0)
    for name, shape, init in entries:
        n = int(np.prod(shape))
        params[name] = flat[o:o + n].view(shape)
        params[name].copy_(spec.init_tensor(shape, init))
        grads[name] = gflat[o:o + n].view(shape)
        o += n
    return params, grads, flat, gflat


EW = {"gn_apply", "gn_backward", "fuse_sum", "upsample_bwd", "grad_add", "conat"}  # element-wise families: they record algorithmic BYTES
TENSOR = {"conv_generic", "conv_k3s1", "conv_pw", "wgrad_generic", "wgrad_k3s1", "wgrad_s2d", "wgrad_pw", "conv_k3s1_units", "wgrad_k3s1_units"}


def ncu_traffic(kname, key, batch):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture) for this
    kernel at this shape, from the newest profiles/r*_ncu_traffic.json that has it; None when never captured."""
    import glob
    want = "%s|%d|%d|%s|b%d" % (kname, key[1], key[2], "x".join(str(v) for v in key[6]), batch)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")), reverse=True):
        try:
            d = json.load(open(path))
        except Exception:
            continue
        e = d.get("by_shape", {}).get(want)
        if e and e.get("traffic_MB"):
            return e["traffic_MB"] * 1e6, os.path.relpath(path, ROOT) + " [" + want + "]"
        if kname in d and key[1] == 32 and key[2] == 32 and tuple(key[6]) == (16, 160, 64) and batch == 16 and d[kname].get("traffic_MB"):
            return d[kname]["traffic_MB"] * 1e6, os.path.relpath(path, ROOT)
    return None, None


def summarize_profile(prof, total_ms, nprof, batch, timing_note):
    """roofline object + top-kernel list from the per-launch CUDA-event pairs ops.PROFILE collected.  The dominant kernel
    is the tensor-kernel FUNCTION with the largest summed time; the figures quoted are those of its most time-consuming
    shape.  `peak` is the measured BURST bf16 peak (the denominator BASELINE.md's >= 60 % target is stated against); the
    fraction of the sustained peak is given beside it."""
    pk_burst, pk_sust, hbm_peak, src = peaks()
    fam, ew = {}, {}
    for key, evs in prof.items():
        if key == "_only":
            continue
        tms = [s.elapsed_time(e) for s, e, _ in evs]
        if key[0] in EW:  # HBM-bound kernels: per function, algorithmic bytes / time vs the measured copy bandwidth
            a = ew.setdefault(key[0], [0.0, 0, 0.0])
            a[0] += sum(tms); a[1] += len(tms); a[2] += sum(f for _, _, f in evs)
            continue
        fam[key] = (sum(tms), len(tms), sum(f for _, _, f in evs))
    hbm_kernels = {k: {"GBps": round(v[2] / (v[0] * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(v[2] / (v[0] * 1e-3) / 1e9 / hbm_peak, 3),
                       "launches_per_step": v[1] // nprof, "ms_per_step": round(v[0] / nprof, 3), "share_of_step": round(v[0] / total_ms, 3)}
                   for k, v in ew.items()}
    roof, top = None, []
    if fam:
        for key, (tt, n, fl) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:16]:
            top.append({"kernel": key[0], "cin": key[1], "cout": key[2], "taps": key[3], "is_os": [key[4], key[5]], "rows": list(key[6]),
                        "launches": n, "ms_total": round(tt, 3), "tflops": round(fl / (tt * 1e-3) / 1e12, 1) if fl > 0 else None})
            if fl <= 0:
                top[-1]["note"] = "restricted to a device-side unit list (the tiles around the targets): no algorithmic FLOPs claimed"
        by_kernel = {}
        for key, (tt, n, fl) in fam.items():
            a = by_kernel.setdefault(key[0], [0.0, 0, 0.0])
            a[0] += tt; a[1] += n; a[2] += fl
        kname = max((k for k in by_kernel if by_kernel[k][2] > 0), key=lambda k: by_kernel[k][0])
        key, (tt, n, fl) = max(((k, v) for k, v in fam.items() if k[0] == kname), key=lambda kv: kv[1][0])
        ach = fl / (tt * 1e-3) / 1e12
        traffic, tsrc = ncu_traffic(kname, key, batch)
        roof = {"bound": "tensor", "kernel": "%s Cin=%d Cout=%d taps=%d" % key[:4], "achieved": ach, "peak": pk_burst,
                "unit": "TFLOP/s", "frac": ach / pk_burst, "frac_of_sustained_peak": ach / pk_sust, "peak_sustained": pk_sust,
                "peak_source": src + " (MEASURED_PEAKS.json bf16_tflops = burst; bf16_tflops_sustained beside it)",
                "algorithmic_flops_per_launch": fl / n, "avg_launch_ms": tt / n, "share_of_step": by_kernel[kname][0] / total_ms,
                "traffic": traffic, "traffic_source": tsrc, "timing": timing_note,
                "hbm_bound_kernels": hbm_kernels, "hbm_peak_GBps": hbm_peak,
                "all_tensor_kernels": {k: {"tflops": round(v[2] / (v[0] * 1e-3) / 1e12, 1) if v[2] > 0 else None,
                                           "frac_of_burst": round(v[2] / (v[0] * 1e-3) / 1e12 / pk_burst, 3) if v[2] > 0 else None,
                                           "ms_per_step": round(v[0] / nprof, 3), "share_of_step": round(v[0] / total_ms, 3)} for k, v in by_kernel.items()}}
    return roof, top


def run_ours(args, rank, world, local_rank):
    from rtpose_b200 import lib, ops, targets
    from rtpose_b200 import dist as rdist
    from rtpose_b200.engine import Engine
    from rtpose_b200.p8 import P8, _stream
    lib.require_device()
    cfg = args.cfg
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, gf_fwd = CFGS[cfg]
    dev = torch.device("cuda", local_rank)
    B = args.batch
    params, grads, flat, gflat = build_params(cfg, dev)
    if world > 1:
        rdist.broadcast_params([flat])  # in place, under no_grad: the views' shared version counter moves with it
    code_w = [1.0] * 45 if reg == 45 else [1.0, 1.5, 2.0]
    eng = Engine(arch, fuse, params, reg, ncls, weight, code_w)

    # ---- synthetic inputs, resident in HBM: raw = a + (b-a)*u, u ~ U[-0.2, 1)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    a, b = norm if norm is not None else (0.0, 1.0)
    D = in_ch
    raw = (a + (b - a) * (torch.rand((B, D) + RAW_SHAPE, device=dev, generator=g) * 1.2 - 0.2)).to(torch.float16)
    rs = np.random.RandomState(99 + rank)
    poses = torch.from_numpy(targets.random_poses(rs, B, GRID)).to(dev)  # random 3-D skeletons, fp64 [B,15,3], resident
    xin = P8(B, D, *GRID, device=dev)
    from rtpose_b200.optim import FlatAdam, one_cycle
    opt = FlatAdam(flat, gflat, wd=0.01, max_norm=35.0) if not args.no_optimizer else None
    it = [0]

    if args.sync_wgrad:
        ops.ASYNC_WGRAD = False
    if args.serial_branches:
        eng.parallel_branches = False

    # N > 1: the flat gradient buffer is all-reduced in 3 slices on a communication stream, each launched as soon as
    # backward has issued the slice's last gradient; the mean comes from seeding backward with 1/world
    sar = None
    if world > 1 and args.overlap_allreduce and not args.late_allreduce:
        from rtpose_b200 import spec as _spec  # noqa: F401
        sar = rdist.SlicedAllReduce(gflat, [(k, v.numel()) for k, v in params.items()], world=world,
                                    fractions=(0.5, 0.35, 0.15)).attach(eng)

    def step_body():
        # the optimizer rewrote the weights: every pack is rebuilt inside the timed region — forked FIRST, so the ~130 small
        # pack launches run beside the ingest / first GroupNorm kernels instead of after them (the first conv joins)
        eng.packs.refresh_async()
        lib.call("rtp_ingest_pack", raw.data_ptr(), B, D, RAW_SHAPE[0], RAW_SHAPE[1], RAW_SHAPE[2], ROI0[0], ROI0[1], ROI0[2],
                 float(a), float(b - a), 1 if norm is not None else 0, xin.struct(), None, _stream())
        tgt = targets.assign_device(poses, GRID, one_hm=(ncls == 1), min_radius=2 if ncls == 1 else 1)
        # the loss gathers the regression output at the target voxels only: the engine evaluates (and back-propagates) the
        # regression branch of the head around them (Engine.forward, reg_targets)
        hm, rg = eng.forward(xin, True, reg_targets=tgt["ind"])
        out = eng.loss(hm, rg, tgt["hm"], tgt["ind"], tgt["mask"], tgt["cat"], tgt["anno_pose"],
                       grad_scale=(1.0 / world) if sar is not None else 1.0)
        eng.backward(grads)
        if sar is not None:
            sar.join()
        elif world > 1:
            rdist.allreduce_flat(gflat, world)
        if opt is not None:
            opt.step_dev()
        return out

    graph = None

    def step():
        if opt is not None:
            lr, mom = one_cycle(it[0], 1000, lr_max=2e-3)  # configs/cruw_pose/hr3d_one_hm_doppler.py:176-179
            opt.set_hyper(lr, mom)
        it[0] += 1
        return graph() if graph is not None else step_body()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
    barrier()
    use_graph = False
    if not args.no_graph:
        # the step is static (same buffers, same launch sequence): capture it once, replay it in the timed region
        from rtpose_b200.graph import StepGraph
        try:
            lib.launch_count = 0
            graph = StepGraph(step_body, warmup=0, high_priority=not os.environ.get("RTP_NO_PRIO")).capture()
            launches_per_step = lib.launch_count
            use_graph = True
            for _ in range(2):  # replays are steps too: keep the schedule moving
                out = step()
        except Exception as ex:  # capture is an optimisation; the eager path is the same work
            graph = None
            print("bench: CUDA-graph capture failed (%r); running eagerly" % (ex,), file=sys.stderr)
    barrier()
    loss0 = float(out[0])
    if not np.isfinite(loss0):
        raise RuntimeError("non-finite loss in warm-up: %r" % loss0)

    # ---- timed region: K steps, CUDA events, max over ranks; L2 (126 MB) is flushed by the working set itself
    #      (one step streams > 10 GB of activations through HBM) — stated in config.l2
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    host_issue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps  # CPU time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    launches = launches_per_step * args.steps if use_graph else lib.launch_count
    ms = e0.elapsed_time(e1) / args.steps
    if args.timeline:
        # diagnostic only (never a bench value): CUPTI kernel records of two replayed steps -> per-kernel (stream, start,
        # duration); tools/timeline_report.py turns them into busy / idle / overlap figures
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        rows = [{"name": e.name, "stream": getattr(e, "stream", None) if hasattr(e, "stream") else None,
                 "ts": e.time_range.start, "dur": e.time_range.end - e.time_range.start} for e in tp.events()
                if str(e.device_type).endswith("CUDA")]
        with open(args.timeline, "w") as f:
            json.dump(rows, f)
        try:
            tp.export_chrome_trace(args.timeline + ".trace.json")
        except Exception as ex:
            print("bench: chrome trace export failed: %r" % (ex,), file=sys.stderr)
    # ---- per-kernel durations for the roofline: the same step, issued eagerly with the weight gradients in-stream so
    #      that every CUDA-event pair brackets exactly one kernel running alone (in the timed region kernels of the
    #      wgrad side stream overlap the main stream, and event timing inside a replayed graph is not available)
    graph, async_was, par_was = None, ops.ASYNC_WGRAD, eng.parallel_branches
    ops.ASYNC_WGRAD, eng.parallel_branches = False, False
    step()
    ops.PROFILE = {"_only": TENSOR | EW}
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nprof = min(args.steps, 5)
    p0.record()
    for _ in range(nprof):
        out = step()
    p1.record()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    ops.ASYNC_WGRAD, eng.parallel_branches = async_was, par_was
    ms_serial = p0.elapsed_time(p1) / nprof
    sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    value = world * B / (ms * 1e-3)

    roof, top = summarize_profile(prof, ms_serial * nprof, nprof, B,
                                  "CUDA events around each launch in %d eager, single-stream steps of the same workload run right "
                                  "after the timed region (%.2f ms/step; the timed region itself replays a CUDA graph with weight "
                                  "gradients on a side stream, %.2f ms/step)" % (nprof, ms_serial, ms))
    pk_burst, pk_sust, hbm_peak, src = peaks()

    line = {"metric": "radar frames/sec HRRadarPose fwd+bwd", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(cfg, B, world, D), **{
                       "l2": "inputs+activations per step >> 126 MB L2 (no explicit flush needed)",
                       "allreduce": (None if world == 1 else ("3 slices overlapped with backward on a communication stream" if sar is not None
                                                              else "one flat all-reduce after backward (overlapped slices measured slower: --overlap-allreduce)")),
                       "cuda_graph": use_graph, "wgrad_side_stream": bool(ops.ASYNC_WGRAD), "branch_streams": bool(eng.parallel_branches),
                       "targets": "assigned on the device from resident fp64 skeletons every step (rtp_assign_targets)",
                       "head_regression_branch": ("training: forward and backward on the (sample, tile) units around the target voxels only — "
                                                  "CenterHead.loss gathers the regression map there and nowhere else, so loss and all "
                                                  "gradients equal the dense evaluation (A/B: RTP_NO_SPARSE_FWD / _UNITS / _REG=1)"),
                       "optimizer": ("fused clip(35) + decoupled wd + Adam, one-cycle lr (rtp_adam_step) inside the timed region"
                                     if opt is not None else "none (--no-optimizer)")}),
            "model_tflops": 3 * gf_fwd * 1e9 * value / 1e12, "model_flops_frac_of_peak": 3 * gf_fwd * 1e9 * value / 1e12 / world / pk_burst,
            "model_flops_frac_of_sustained_peak": 3 * gf_fwd * 1e9 * value / 1e12 / world / pk_sust,
            "model_tflops_is": "3 x the reference model's dense forward FLOPs x frames/s (MFU convention: the work the reference performs per frame; "
                               "the sparse regression branch executes fewer tensor FLOPs than that)",
            "loss": float(out[0]), "gpu_launches": launches, "host_issue_ms_per_step": round(host_issue_ms, 2), "clocks": sampler.summary(), "roofline": roof, "top_kernels": top}

    if not args.no_extras:
        line["e2e"] = e2e_public_api(args, dev, rank, world)  # every rank takes part (gradient all-reduce inside)
    if rank == 0 and world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_baseline(args)
        try:
            line["inference"] = inference_bench(args, eng, dev)
        except Exception as ex:  # extras must never take the headline down
            line["inference"] = {"error": repr(ex)}
        try:
            line["loader"] = loader_bench(args, dev)
        except Exception as ex:
            line["loader"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line))


def run_dcn_head(args, rank, world, local_rank):
    """BASELINE.json configs[4]: the wide-channel (phase) HRRadarPose variant with the deformable head, data-parallel
    training.  The head is the 3-D-compatible DCNSepHead (det3d_compat, dcn_head='fold_z': 2-D ops on the z-folded batch),
    a module-level composition, so this leg drives the det3d-style model — model(example) -> loss -> backward -> flat
    gradient all-reduce -> fused clip + Adam — eagerly (no CUDA graph), inputs resident in HBM."""
    from rtpose_b200 import det3d_compat as D
    from rtpose_b200 import dist as rdist
    from rtpose_b200 import lib, ops, targets
    from rtpose_b200.optim import FlatAdam, one_cycle
    lib.require_device()
    cfg = args.cfg
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, gf_fwd = CFGS[cfg]
    dev = torch.device("cuda", local_rank)
    B = args.batch
    names = ["Pelvis"]
    head_in = fout if fuse != "top" else 32
    model_cfg = dict(type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
                     backbone=dict(type="HRNet3D", backbone_cfg=arch, final_conv_in=fin, final_conv_out=fout, final_fuse=fuse, ds_factor=1),
                     pose_head=dict(type="CenterHead", tasks=[dict(num_class=ncls, class_names=names[:ncls])], in_channels=head_in,
                                    share_conv_channel=head_in, dataset="cruw_pose", weight=weight,
                                    code_weights=[1.0] * 45, common_heads={"reg": (reg, 2)}, dcn_head="fold_z"),
                     neck=None)
    torch.manual_seed(0)
    model = D.build_detector(model_cfg, train_cfg=None, test_cfg=None).to(dev)
    model.pose_head.sync_free_losses = True
    params = [p for p in model.parameters()]
    total = sum(p.numel() for p in params)
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    gflat = torch.zeros(total, dtype=torch.float32, device=dev)
    o = 0
    for p in params:  # parameters become views of one flat buffer (fused optimizer, one all-reduce)
        flat[o:o + p.numel()].copy_(p.data.reshape(-1))
        p.data = flat[o:o + p.numel()].view(p.shape)
        o += p.numel()
    if world > 1:
        rdist.broadcast_params([flat])
    opt = FlatAdam(flat, gflat, wd=0.01, max_norm=35.0, params=params)
    rs = np.random.RandomState(7 + rank)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.clamp(torch.rand((B, in_ch) + GRID, device=dev, generator=g) * 1.2 - 0.2, min=0)
    tg = targets.assign(targets.random_poses(rs, B, GRID), GRID, one_hm=True, min_radius=2)
    ex = {"rdr": {"rdr_tensor": x}, "meta": [{}] * B}
    for k, v in tg.items():
        ex["rdr"][k] = [torch.from_numpy(v).to(dev)]
    it = [0]

    def step():
        lr, mom = one_cycle(it[0], 1000, lr_max=2e-3)
        it[0] += 1
        for p in params:
            p.grad = None
        losses = model(ex, return_loss=True)
        losses["loss"][0].backward()
        torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params], out=gflat)
        if world > 1:
            rdist.allreduce_flat(gflat, world)
        opt.step(lr, mom)
        return losses["loss"][0]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        loss = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    launches = lib.launch_count
    ms = e0.elapsed_time(e1) / args.steps
    ops.PROFILE = {"_only": TENSOR | EW}
    ops.ASYNC_WGRAD = False
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step()
    p1.record()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    sampler.stop_flag = True
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    roof, top = summarize_profile(prof, p0.elapsed_time(p1), 1, B, "CUDA events around each conv / element-wise launch of one eager step "
                                  "run after the timed region (DCN sampling / scatter kernels are not itemised)")
    value = world * B / (ms * 1e-3)
    line = {"metric": "radar frames/sec HRRadarPose fwd+bwd", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s + deformable head (CenterHead dcn_head='fold_z': FeatureAdaption x2 at C=%d, DCN v1 dg=4 on the "
                                   "z-folded batch [B*16, %d, 64, 160]) training fwd+bwd, batch %d per GPU, model input resident in HBM "
                                   "(BASELINE.json configs[4])" % (cfg, head_in, head_in, B),
                       "per_gpu_batch": B, "global_batch": B * world, "grid": list(GRID), "parallelism": "dp%d" % world,
                       "cuda_graph": False, "optimizer": "fused clip(35) + decoupled wd + Adam (rtp_adam_step) inside the timed region",
                       "l2": "activations per step >> 126 MB L2"},
            "loss": float(loss), "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "top_kernels": top,
            "params": total}
    if rank == 0:
        print(json.dumps(line))


def e2e_public_api(args, dev, rank=0, world=1):
    """Same metric through the reference-facing API (build_detector -> model(example) -> loss.backward()) with HOST
    buffers: per step the fp32 rdr_tensor + targets go pinned-host -> device and the loss comes back.  At N > 1 every
    rank steps its own shard and the gradients are averaged after backward() the way the reference's trainer does
    (rtpose_b200.dist.allreduce_grads <-> det3d/core/utils/dist_utils.py:40-57); time = max over ranks."""
    from rtpose_b200 import det3d_compat as D
    from rtpose_b200 import targets
    from rtpose_b200 import dist as rdist
    cfg = args.cfg
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, _ = CFGS[cfg]
    B = args.batch
    names = ["Pelvis", "Right_Hip", "Right_Knee", "Right_Ankle", "Left_Hip", "Left_Knee", "Left_Ankle", "Thomx", "Head",
             "Left_Shoulder", "Left_Elbow", "Left_Wrist", "Right_Shoulder", "Right_Elbow", "Right_Wrist"]
    head_in = fout if fuse != "top" else 32
    model_cfg = dict(type="RadarPoseNet", pretrained=None, reader=dict(type="RadarFeatureNet"),
                     backbone=dict(type="HRNet3D", backbone_cfg=arch, final_conv_in=fin, final_conv_out=fout, final_fuse=fuse, ds_factor=1),
                     pose_head=dict(type="CenterHead", tasks=[dict(num_class=ncls, class_names=names[:ncls])], in_channels=head_in,
                                    share_conv_channel=head_in, dataset="cruw_pose", weight=weight,
                                    code_weights=[1.0] * 45 if reg == 45 else [1.0, 1.5, 2.0], common_heads={"reg": (reg, 2)}, dcn_head=False),
                     neck=None)
    torch.manual_seed(0)
    model = D.build_detector(model_cfg, train_cfg=None, test_cfg=None).to(dev)
    model.pose_head.sync_free_losses = True
    model.cuda_graph = not args.no_graph  # one replayed graph per training step (det3d_compat._StepJob)
    if world > 1:
        rdist.broadcast_params(list(model.parameters()))
    rs = np.random.RandomState(7 + rank)
    x_host = torch.from_numpy(np.maximum(rs.uniform(-0.2, 1.0, (B, in_ch) + GRID), 0).astype(np.float32)).pin_memory()
    tg = targets.assign(targets.random_poses(rs, B, GRID), GRID, one_hm=(ncls == 1), min_radius=2 if ncls == 1 else 1)
    t_host = {k: torch.from_numpy(v).pin_memory() for k, v in tg.items()}
    h2d = x_host.numel() * 4 + sum(v.numel() * v.element_size() for v in t_host.values())

    # Every step copies ITS OWN inputs host -> device; the copy of step i+1 is issued on a side stream while step i
    # computes (what a prefetching loader does; the reference's trainer has an unused Prefetcher for this, trainer.py:119-140)
    copy_stream = torch.cuda.Stream(device=dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            ex = {"rdr": {"rdr_tensor": x_host.to(dev, non_blocking=True)}, "meta": [{}] * B}
            for k, v in t_host.items():
                ex["rdr"][k] = [v.to(dev, non_blocking=True)]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ex, ev

    # The loss of every step is read back to the host (4 bytes, pinned) — asynchronously: the copy is queued behind the
    # step and the VALUE is consumed one step later, after the next step has been enqueued, so the host never stalls the
    # device between steps (what a trainer that logs the loss does with a one-step lag).  Every step's loss is read inside
    # the timed region; the last one after the loop.
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending = []   # (slot, event) of losses in flight
    seen = []

    def drain(keep):
        while len(pending) > keep:
            slot, evl = pending.pop(0)
            evl.synchronize()
            seen.append(float(loss_host[slot]))

    def step(ex, ev, i):
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        ex["rdr"]["rdr_tensor"].record_stream(cur)
        for k in t_host:
            ex["rdr"][k][0].record_stream(cur)
        for p in model.parameters():
            p.grad = None
        losses = model(ex, return_loss=True)
        losses["loss"][0].backward()
        if world > 1:
            rdist.allreduce_grads(params, world)
        slot = i & 1
        loss_host[slot:slot + 1].copy_(losses["loss"][0].detach().reshape(1), non_blocking=True)  # D2H read of the step's result
        evl = torch.cuda.Event()
        evl.record(cur)
        pending.append((slot, evl))
        nxt = fetch()                                        # overlaps with the GPU work just enqueued
        drain(1)                                             # consume the PREVIOUS step's loss
        return nxt

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    params = list(model.parameters())
    nxt = fetch()
    it = 0
    for _ in range(max(3, min(args.warmup, 3))):
        nxt = step(*nxt, it)
        it += 1
    drain(0)
    barrier()
    steps = max(3, min(args.steps, 10))
    seen.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        nxt = step(*nxt, it)
        it += 1
    drain(0)
    e1.record()
    barrier()
    assert len(seen) == steps and all(np.isfinite(v) for v in seen), seen
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    return {"value": world * B / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
            "bytes_are": "per rank", "ms_per_step": ms, "cuda_graph": bool(model.cuda_graph),
            "loss_readback": "every step, asynchronous: consumed on the host one step later (no host stall between steps)",
            "api": "det3d_compat.build_detector(...)(example, return_loss=True); loss.backward()"}


def cpu_baseline(args):
    """The reference's modules (baseline/_ref; the oracle port when absent) on the box's host cores: bounded samples at
    batch 1, full grid — training fwd+bwd (the headline metric) and forward + decode (BASELINE.json configs[0])."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tstep, istep, kind = reference_runner(args.cfg, 1)
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        tstep()
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times[1:]))
    itimes = []
    for i in range(4):
        t0 = time.perf_counter()
        istep()
        itimes.append(time.perf_counter() - t0)
    di = float(np.median(itimes[1:]))
    return {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "batch 1 fwd+bwd at 16x64x160, fp32, 1 warm-up + median of 3 (%d threads)" % cores,
            "fwd_decode_b1": {"value": 1.0 / di, "unit": "frames/s", "ms": di * 1e3,
                              "sample": "configs[0]: batch 1 forward + predict/decode on the host cores, median of 3"}}


def loader_bench(args, dev):
    """on-disk fp16 cubes -> P8 model input through rtpose_b200.loader.CubeLoader (ROI-row preads ->
    pinned slab -> H2D -> rtp_ingest_pack), beside the reference's way of producing the same tensor in one process
    (np.load + astype(float32) + crop + normalise + clamp, cruw_pose.py:170-185; torch.tensor + blocking H2D,
    cruw_pose.py:264-266, train.py(apis):27-62).  Files are synthetic, written to a temp dir and read from the page cache."""
    import shutil
    import tempfile
    from rtpose_b200 import loader
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, _ = CFGS[args.cfg]
    B, nfiles = args.batch, 2 * args.batch
    shape = ((in_ch,) if in_ch > 1 else ()) + RAW_SHAPE
    a, b = norm if norm is not None else (0.0, 1.0)
    need = 2 * nfiles * int(np.prod(shape)) * 2
    shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need
    tmp = tempfile.mkdtemp(prefix="rtp_cubes_", dir="/dev/shm" if shm_ok else None)
    try:
        rs = np.random.RandomState(11)
        base = (a + (b - a) * (rs.rand(*shape).astype(np.float32) * 1.2 - 0.2)).astype(np.float16)
        paths = []
        for i in range(nfiles):
            paths.append(os.path.join(tmp, "%06d.npy" % i))
            np.save(paths[-1], np.roll(base, i, axis=-1))
        ld = loader.CubeLoader(paths * 4, batch=B, norm=norm, depth=3)  # 8 batches per epoch over 32 distinct files
        for _ in ld:  # warm-up epoch (page cache, pinned allocations)
            pass
        torch.cuda.synchronize()
        epochs, t0 = 3, time.perf_counter()
        for _ in range(epochs):
            for x, _p in ld:
                pass
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ours = epochs * len(ld) * B / dt
        # the reference's way, same files, one process (it uses 2 DataLoader workers: build_loader.py:46-57)
        z0, y0, x0 = ROI0
        t0 = time.perf_counter()
        for p in paths[:B]:
            arr = np.load(p).astype(np.float32)
            arr = arr[..., z0:z0 + GRID[0], y0:y0 + GRID[1], x0:x0 + GRID[2]]
            arr = (arr - a) / (b - a)
            arr[arr < 0.] = 0.
            t = torch.tensor(arr[None]).to(dev)
        torch.cuda.synchronize()
        ref = B / (time.perf_counter() - t0)
        return {"value": ours, "unit": "frames/s", "reference_style_1proc": ref, "file_MB": round(base.nbytes / 1e6, 1),
                "read_MB_per_frame": round(ld.bytes_per_frame / 1e6, 1), "files": nfiles, "batches_per_epoch": len(ld), "batch": B,
                "source": "page cache (%s)" % os.path.dirname(paths[0]), "includes": "pread ROI rows + H2D + ingest kernel"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def inference_bench(args, eng, dev):
    """BASELINE.json configs[1]: inference, batch 32, bf16, incl. heat-map arg-max keypoint decode — a second leg of the
    same line with its own roofline (per-launch CUDA events of an eager single-stream pass, like the training leg)."""
    from rtpose_b200 import lib, ops
    from rtpose_b200.p8 import P8, _stream
    arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, gf_fwd = CFGS[args.cfg]
    B = 32
    g = torch.Generator(device=dev).manual_seed(5)
    a, b = norm if norm is not None else (0.0, 1.0)
    raw = (a + (b - a) * (torch.rand((B, in_ch) + RAW_SHAPE, device=dev, generator=g) * 1.2 - 0.2)).to(torch.float16)
    xin = P8(B, in_ch, *GRID, device=dev)

    def step():
        lib.call("rtp_ingest_pack", raw.data_ptr(), B, in_ch, RAW_SHAPE[0], RAW_SHAPE[1], RAW_SHAPE[2], ROI0[0], ROI0[1], ROI0[2],
                 float(a), float(b - a), 1 if norm is not None else 0, xin.struct(), None, _stream())
        hm, rg = eng.forward(xin, False)
        return eng.decode(hm, rg, (0.0453125, 0.15703125, 0.3625), (0.7703125, -5.025, -1.0875))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    run, graphed = step, False
    lib.launch_count = 0
    step()
    launches = lib.launch_count
    if not args.no_graph:  # the inference step is static too: replay it as one graph (weights packed once, outside)
        from rtpose_b200.graph import StepGraph
        try:
            run = StepGraph(step, warmup=0).capture()
            graphed = True
            run()
        except Exception as ex:
            run = step
            print("bench: inference graph capture failed (%r); running eagerly" % (ex,), file=sys.stderr)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        idx, score, xyz = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # per-kernel pass: eager, one stream
    par_was = eng.parallel_branches
    eng.parallel_branches = False
    step()
    ops.PROFILE = {"_only": TENSOR | EW}
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(3):
        step()
    p1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    eng.parallel_branches = par_was
    ms_serial = p0.elapsed_time(p1) / 3
    roof, top = summarize_profile(prof, ms_serial * 3, 3, B, "CUDA events around each launch in 3 eager single-stream inference steps "
                                  "(%.2f ms/step; the timed region replays a CUDA graph, %.2f ms/step)" % (ms_serial, ms))
    pk_burst, pk_sust, _, _ = peaks()
    val = B / (ms * 1e-3)
    return {"metric": "radar frames/sec HRRadarPose inference incl. decode", "value": val, "unit": "frames/s", "batch": B,
            "ms_per_step": ms, "steps": n, "cuda_graph": graphed, "dtype": "bf16",
            "config": {"workload": "%s inference, batch %d, raw fp16 cube -> ingest -> HRNet3D -> CenterHead -> arg-max keypoint decode "
                                   "(BASELINE.json configs[1])" % (args.cfg, B)},
            "model_tflops": gf_fwd * 1e9 * val / 1e12, "model_flops_frac_of_peak": gf_fwd * 1e9 * val / 1e12 / pk_burst,
            "gpu_launches_per_step": launches, "roofline": roof, "top_kernels": top[:8]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="hr3d_one_hm_doppler", choices=sorted(CFGS))
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--serial-branches", action="store_true", help="run the HR-module branches one after the other on one stream")
    ap.add_argument("--sync-wgrad", action="store_true", help="run weight gradients in-stream (no side stream)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / inference legs")
    ap.add_argument("--timeline", default=None, help="diagnostic: write the CUPTI kernel timeline of two replayed steps to this JSON file")
    ap.add_argument("--no-optimizer", action="store_true", help="time forward+backward only (no fused clip+Adam step)")
    ap.add_argument("--dcn-head", action="store_true", help="configs[4]: the deformable head (dcn_head='fold_z') on the det3d-style model")
    ap.add_argument("--late-allreduce", action="store_true", help="N > 1: one all-reduce after backward (the default since the end of round 2; kept for old command lines)")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="N > 1: all-reduce in 3 slices on a communication stream while backward runs (A/B: measured SLOWER, 20.24 vs 19.31 ms at N = 2 — "
                         "the NCCL kernels hold SMs the persistent one-CTA-per-SM conv kernels need)")
    ap.add_argument("--ref-gpu-probe", action="store_true", help="time the unmodified reference on this GPU through stock PyTorch (informative)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.ref_gpu_probe:
        if rank == 0:
            ref_gpu_probe(args)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.dcn_head:
            run_dcn_head(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
