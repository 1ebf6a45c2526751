"""Ingest kernel at the bench shape: raw fp16 [16, 32, 32, 128, 256] -> ROI -> normalise -> P8 bf16."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtpose_b200 import lib  # noqa: E402
from rtpose_b200.p8 import P8, _stream  # noqa: E402

B, D = 16, 32
raw = (torch.rand(B, D, 32, 128, 256, device="cuda") * 12 - 2).half()
x = P8(B, D, 16, 64, 160)


def run():
    lib.call("rtp_ingest_pack", raw.data_ptr(), B, D, 32, 128, 256, 13, 32, 17, 0.0, 10.0, 1, x.struct(), None, _stream())


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
mb = B * D * 16 * 64 * 160 * 2 * 2 / 1e6
print("ingest                           %.3f ms  %7.1f GB/s (algorithmic %d MB)" % (ms, mb / ms, mb))
