"""Bound analysis of wgrad_k3s1: time the kernel with X copies / dY copies / MMAs disabled (debug switch)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtpose_b200 import lib, ops
from rtpose_b200.p8 import P8

L = lib.load()
flag = C.c_int.in_dll(L, "rtp_wgrad_k3s1_dbg")
N, Cc, Z, Y, X = 16, 32, 16, 64, 160
x = P8.from_ncdhw(torch.randn(N, Cc, Z, Y, X, device="cuda"))
dy = P8.from_ncdhw(torch.randn(N, Cc, Z, Y, X, device="cuda"))
dW = torch.zeros(32, 32, 3, 3, 3, device="cuda")
for mode, name in ((0, "full"), (1, "no X copies"), (4, "no dY copies"), (5, "no copies (MMA issue only)"), (2, "no MMAs (loads only)"),
                   (7, "nothing (protocol only)")):
    flag.value = mode
    for _ in range(3):
        ops.conv_wgrad(x, dy, 3, 1, dW)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.conv_wgrad(x, dy, 3, 1, dW)
    b.record()
    torch.cuda.synchronize()
    print("%-32s %.3f ms" % (name, a.elapsed_time(b) / 20))
flag.value = 0
