"""Sweep of CubeLoader reader parallelism on the box (files in /dev/shm, page-cache reads)."""
import os, shutil, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtpose_b200 import loader

B, n = 16, 32
tmp = tempfile.mkdtemp(prefix="rtp_cubes_", dir="/dev/shm")
try:
    base = (np.random.RandomState(0).rand(32, 32, 128, 256) * 12 - 2).astype(np.float16)
    paths = []
    for i in range(n):
        paths.append(os.path.join(tmp, "%06d.npy" % i)); np.save(paths[-1], base)
    # raw reader speed, no GPU involved
    out = torch.empty((B, 32, 16, 64, 256), dtype=torch.float16).pin_memory()
    for th in (1, 2, 4, 8, 16):
        t0 = time.perf_counter()
        for r in range(3):
            for i in range(B):
                loader.read_roi_slab(paths[i], out=out[i], threads=th)
        dt = (time.perf_counter() - t0) / (3 * B)
        print("read_roi_slab threads=%2d: %.2f ms/frame  %.1f GB/s" % (th, dt * 1e3, 16.8e-3 / dt))
    for fw, io in ((1, 8), (1, 16), (2, 4), (2, 8), (4, 4), (8, 2), (16, 1)):
        ld = loader.CubeLoader(paths * 4, batch=B, norm=(0.0, 10.0), depth=3, frame_workers=fw, io_threads=io)
        for _ in ld: pass
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4):
            for x, _p in ld: pass
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("CubeLoader frame_workers=%2d io_threads=%d: %.0f frames/s" % (fw, io, 4 * len(ld) * B / dt))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
