"""Host-side proof of the index arithmetic the trilinear-upsample transpose kernels rely on (csrc/fuse.cu: ac_axis,
hat_weight, hat_range, the shared-memory row bound of upsample_bwd_yx_kernel), restated in numpy float32 — the same
operations in the same precision as the device code — and checked against ATen's align_corners interpolation
(F.interpolate, the reference's call at hr_util/hr3d.py:219 and backbones/hrnet3d.py:39)."""
import math

import numpy as np
import pytest
import torch

f32 = np.float32
PAIRS = [(80, 160), (40, 160), (20, 160), (32, 64), (16, 64), (8, 64), (8, 16), (4, 16), (2, 16), (2, 3), (3, 5), (5, 10), (7, 14),
         (3, 6), (2, 4), (4, 7), (20, 40), (10, 40), (40, 80), (5, 7), (2, 160), (1, 6)]


def ac_scale(n_in, n_out):
    return f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)


def hat_weight(d, l, scale):
    src = scale * f32(d)
    w = f32(1) - (src - f32(l)) if src >= f32(l) else src - f32(l - 1)
    return w if w > 0 else f32(0)


def hat_range(l, n_hi, scale):
    inv = f32(1) / scale
    lo = max(0, int(math.floor(f32(l - 1) * inv)))
    hi = min(n_hi - 1, int(math.ceil(f32(l + 1) * inv)))
    if scale * f32(lo) <= f32(l - 1):
        lo += 1
    if scale * f32(hi) >= f32(l + 1):
        hi -= 1
    return lo, hi


def forward_matrix(n_lo, n_hi):
    """U[d, l]: weight of low-res source l in high-res destination d, from ATen itself."""
    eye = torch.eye(n_lo, dtype=torch.float32).view(n_lo, 1, n_lo, 1, 1)  # batch = one-hot source, interpolate along z
    up = torch.nn.functional.interpolate(eye, size=(n_hi, 1, 1), mode="trilinear", align_corners=True)
    return up.view(n_lo, n_hi).t().numpy()


@pytest.mark.parametrize("n_lo,n_hi", PAIRS)
def test_hat_weights_are_the_transposed_interpolation(n_lo, n_hi):
    U = forward_matrix(n_lo, n_hi)
    scale = ac_scale(n_lo, n_hi)
    for l in range(n_lo):
        for d in range(n_hi):
            w = hat_weight(d, l, scale) if scale > 0 else f32(1)
            assert abs(float(w) - float(U[d, l])) <= 2e-6, (n_lo, n_hi, d, l, w, U[d, l])


@pytest.mark.parametrize("n_lo,n_hi", [p for p in PAIRS if p[0] > 1])
def test_candidate_range_is_complete_and_tile_bound_holds(n_lo, n_hi):
    scale = ac_scale(n_lo, n_hi)
    ranges = []
    for l in range(n_lo):
        lo, hi = hat_range(l, n_hi, scale)
        nz = [d for d in range(n_hi) if hat_weight(d, l, scale) > 0]
        assert nz and lo <= nz[0] and nz[-1] <= hi, (l, lo, hi, nz)           # nothing with weight is left out
        assert lo >= nz[0] - 1 and hi <= nz[-1] + 1                          # and at most one zero-weight candidate per end
        ranges.append((lo, hi))
    assert all(a[0] <= b[0] and a[1] <= b[1] for a, b in zip(ranges, ranges[1:]))  # monotone: a tile's rows are one interval
    kTXL = 8
    nrows_max = int(math.ceil(f32(kTXL + 1) / scale)) + 4  # rtp_upsample_bwd's shared-memory bound
    for l0 in range(0, n_lo, kTXL):
        l1 = min(n_lo, l0 + kTXL) - 1
        assert ranges[l1][1] - ranges[l0][0] + 1 <= nrows_max - 1, (l0, l1, ranges[l0], ranges[l1], nrows_max)


@pytest.mark.parametrize("n_lo,B", [(32, 2), (16, 4), (8, 8), (16, 2), (8, 4), (8, 2), (4, 4), (4, 2), (4, 8), (32, 4), (32, 8)])
def test_hat_support_lies_in_the_blocks_of_the_neighbouring_lanes(n_lo, B):
    """upsample_bwd_yx_shfl_kernel (Y == B * Yl): lane l holds the B high-resolution vectors [l*B, (l+1)*B) of a row and
    evaluates the hat of source index l from its own block and the blocks of lanes l-1 and l+1 — every destination with a
    non-zero weight must lie in [(l-1)*B, (l+2)*B), and visiting those candidates in ascending order with the hat weight
    itself (zero outside the hat) reproduces the exact candidate range of the gather kernel."""
    n_hi = n_lo * B
    scale = ac_scale(n_lo, n_hi)
    for l in range(n_lo):
        nz = [d for d in range(n_hi) if hat_weight(d, l, scale) > 0]
        assert nz[0] >= (l - 1) * B and nz[-1] < (l + 2) * B, (n_lo, B, l, nz)
        lo, hi = hat_range(l, n_hi, scale)
        cand = [d for d in range(max(0, (l - 1) * B), min(n_hi, (l + 2) * B))]
        assert [d for d in cand if hat_weight(d, l, scale) > 0] == nz
        assert all(hat_weight(d, l, scale) == 0 for d in cand if d < lo or d > hi)
