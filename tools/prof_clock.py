"""Is the dominant conv clock-limited in a sustained run?  Times rtp_conv_k3s1 32->32 (bench shape) launch by launch while
sampling the SM clock / power through NVML, first after an idle period (boost clock) and then in a sustained loop."""
import os, sys, time, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import ops  # noqa: E402
from rtpose_b200.p8 import P8  # noqa: E402
import pynvml as N  # noqa: E402

N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], [False]


def sampler():
    while not stop[0]:
        samples.append((time.perf_counter(), N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        N.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.002)


B, C, grid = 16, 32, (16, 64, 160)
x = P8.from_ncdhw(torch.randn(B, C, *grid, device="cuda"))
w = torch.randn(C, C, 3, 3, 3, device="cuda") * 0.05
out = P8(B, C, *grid)
packs = ops.PackedWeights()
ops.conv_forward(packs, x, w, 1, out)
torch.cuda.synchronize()
flops = 2.0 * B * 163840 * C * C * 27
th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(1.0)
n = 400
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
t0 = time.perf_counter()
evs[0].record()
for i in range(n):
    ops.conv_forward(packs, x, w, 1, out)
    evs[i + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
stop[0] = True
th.join()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
print("launch time (ms): first 5 %s | #20-25 %s | last 5 %s" % ([round(v, 3) for v in ms[:5]], [round(v, 3) for v in ms[20:25]],
                                                                  [round(v, 3) for v in ms[-5:]]))
print("TFLOP/s: first launch %.0f, median of the last 100 %.0f" % (flops / ms[0] / 1e9, flops / sorted(ms[-100:])[50] / 1e9))
idle = [s for s in samples if s[0] < t0 - 0.2]
busy = [s for s in samples if t0 + 0.02 < s[0] < t1]
print("idle : SM %d MHz, %.0f W" % (sorted(s[1] for s in idle)[len(idle) // 2], sorted(s[2] for s in idle)[len(idle) // 2]))
if busy:
    print("busy : SM clock MHz min %d median %d max %d; power W median %.0f max %.0f; reasons mask OR %#x; %d samples over %.0f ms" %
          (min(s[1] for s in busy), sorted(s[1] for s in busy)[len(busy) // 2], max(s[1] for s in busy),
           sorted(s[2] for s in busy)[len(busy) // 2], max(s[2] for s in busy), __import__("functools").reduce(lambda a, b: a | b, [s[3] for s in busy]),
           len(busy), (t1 - t0) * 1e3))
