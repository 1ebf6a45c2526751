"""Experiment: the bench step (ingest -> fwd -> loss -> bwd) as L independent half-batch "lanes" on L streams inside one
CUDA graph, to measure how much element-wise time hides under the other lane's tensor kernels.
    python tools/exp_lanes.py [B] [lanes...]"""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rtpose_b200 import lib, ops, targets  # noqa: E402
from rtpose_b200.engine import Engine  # noqa: E402
from rtpose_b200.graph import StepGraph  # noqa: E402
from rtpose_b200.p8 import P8, _stream  # noqa: E402

cfg = os.environ.get("CFG", "hr3d_one_hm_doppler")
arch, fin, fout, fuse, reg, ncls, weight, in_ch, norm, gf = bench.CFGS[cfg]
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lanes_list = [int(v) for v in sys.argv[2:]] or [1, 2]
a, b = norm if norm is not None else (0.0, 1.0)
g = torch.Generator(device=dev).manual_seed(1234)
raw = (a + (b - a) * (torch.rand((B, in_ch) + bench.RAW_SHAPE, device=dev, generator=g) * 1.2 - 0.2)).to(torch.float16)
rs = np.random.RandomState(99)
poses = torch.from_numpy(targets.random_poses(rs, B, bench.GRID)).to(dev)
code_w = [1.0] * 45 if reg == 45 else [1.0, 1.5, 2.0]
for L in lanes_list:
    nb = B // L
    lanes = []
    for i in range(L):
        params, grads, flat, gflat = bench.build_params(cfg, dev)
        eng = Engine(arch, fuse, params, reg, ncls, weight, code_w)
        lanes.append((eng, grads, P8(nb, in_ch, *bench.GRID, device=dev), raw[i * nb:(i + 1) * nb], poses[i * nb:(i + 1) * nb],
                      torch.cuda.Stream(device=dev)))

    def lane_body(i):
        eng, grads, xin, rw, ps, st = lanes[i]
        ops.LANE = i
        lib.call("rtp_ingest_pack", rw.data_ptr(), nb, in_ch, *bench.RAW_SHAPE, *bench.ROI0, float(a), float(b - a),
                 1 if norm is not None else 0, xin.struct(), None, _stream())
        tgt = targets.assign_device(ps, bench.GRID, one_hm=(ncls == 1), min_radius=2 if ncls == 1 else 1)
        if os.environ.get("FWD_ONLY"):  # forward (with tape) + loss only: how much of the FORWARD pass hides under the other lane
            hm, rg = eng.forward(xin, True)
            out = eng.loss(hm, rg, tgt["hm"], tgt["ind"], tgt["mask"], tgt["cat"], tgt["anno_pose"])
            eng.tape = []
            ops.LANE = 0
            return out
        hm, rg = eng.forward(xin, True)
        out = eng.loss(hm, rg, tgt["hm"], tgt["ind"], tgt["mask"], tgt["cat"], tgt["anno_pose"])
        eng.backward(grads)
        ops.LANE = 0
        return out

    def body():
        main = torch.cuda.current_stream()
        for eng, *_ in lanes:
            eng.packs.refresh_async()
        outs = []
        for i in range(1, L):
            lanes[i][5].wait_stream(main)
        for i in range(L):
            if i == 0:
                outs.append(lane_body(0))
            else:
                with torch.cuda.stream(lanes[i][5]):
                    outs.append(lane_body(i))
        for i in range(1, L):
            main.wait_stream(lanes[i][5])
        return outs

    for _ in range(3):
        body()
    torch.cuda.synchronize()
    gr = StepGraph(body, warmup=0, high_priority=(os.environ.get("PRIO", "1") == "1")).capture()
    for _ in range(3):
        gr()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("cfg %s batch %d lanes %d: %.3f ms/step  %.1f frames/s" % (cfg, B, L, ms, B / ms * 1e3), flush=True)
    del lanes, gr
    torch.cuda.empty_cache()
