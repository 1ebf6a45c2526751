"""Host-side proof that the (row grid, tap list, stride, offset) descriptors ops.py builds for rtp_conv / rtp_conv_multi
express conv3d and its input gradient: the descriptor semantics documented in include/rtpose_b200.h
("for tap t the A row is the input vector at (rz*IS+tz, rx*IS+tx, ry*IS+ty) (zero outside), the B tile is packed-weight tap
wt[t]; the result goes to output voxel (rz*OS+oz0, ...)") are emulated in float64 numpy and compared with torch's
conv3d / conv_transpose3d for the four uses the engine makes of them (forward stride 1 / 2, dgrad stride 1, dgrad stride 2 by
parity class), on even and odd extents.  P8 axes: z = tensor dim 2, y = dim 3 (H), x = dim 4 (W)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rtpose_b200 import ops


def emulate(inp, wtaps, out_shape, taps, rows, IS=1, OS=1, off=(0, 0, 0)):
    """inp [N,K,Z,Y,X]; wtaps [ntaps_total][K][Nout] (tap index = wt); returns out [N,Nout,*out_shape(Z,Y,X)]."""
    N, K, Z, Y, X = inp.shape
    out = np.zeros((N, wtaps.shape[2]) + tuple(out_shape))
    RZ, RX, RY = rows
    for rz in range(RZ):
        for rx in range(RX):
            for ry in range(RY):
                acc = np.zeros((N, wtaps.shape[2]))
                for tz, tx, ty, wt in taps:
                    z, x, y = rz * IS + tz, rx * IS + tx, ry * IS + ty
                    if 0 <= z < Z and 0 <= x < X and 0 <= y < Y:
                        acc += inp[:, :, z, y, x] @ wtaps[wt]
                out[:, :, rz * OS + off[0], ry * OS + off[2], rx * OS + off[1]] = acc
    return out


def packs(w):
    """mode 0 (forward): [tap][ci][co]; mode 1 (dgrad): [tap][co][ci]; tap = (kz*k + ky)*k + kx (rtp_weight_pack)."""
    co, ci = w.shape[:2]
    flat = w.reshape(co, ci, -1)
    return flat.transpose(2, 1, 0).copy(), flat.transpose(2, 0, 1).copy()


@pytest.mark.parametrize("grid", [(4, 6, 8), (5, 7, 9), (2, 3, 4)])
def test_forward_and_dgrad_descriptors(grid):
    rs = np.random.RandomState(sum(grid))
    Z, Y, X = grid
    x = rs.randn(2, 3, Z, Y, X)
    w = rs.randn(4, 3, 3, 3, 3)
    fwd, dg = packs(w)
    xt, wt = torch.from_numpy(x), torch.from_numpy(w)
    # forward, stride 1: rows = output grid = input grid
    got = emulate(x, fwd, grid, ops.taps_fwd(3), (Z, X, Y))
    np.testing.assert_allclose(got, F.conv3d(xt, wt, padding=1).numpy(), atol=1e-12)
    # forward, stride 2 (pad 1): rows = output grid, IS = 2
    oz, oy, ox = (Z + 1) // 2, (Y + 1) // 2, (X + 1) // 2
    got = emulate(x, fwd, (oz, oy, ox), ops.taps_fwd(3), (oz, ox, oy), IS=2)
    ref2 = F.conv3d(xt, wt, stride=2, padding=1)
    assert ref2.shape[2:] == (oz, oy, ox)
    np.testing.assert_allclose(got, ref2.numpy(), atol=1e-12)
    # dgrad, stride 1: rows = dx grid, operand dy
    dy = rs.randn(2, 4, Z, Y, X)
    got = emulate(dy, dg, grid, ops.taps_dgrad_s1(3), (Z, X, Y))
    np.testing.assert_allclose(got, F.conv_transpose3d(torch.from_numpy(dy), wt, padding=1).numpy(), atol=1e-12)
    # dgrad, stride 2: one class per parity of the dx voxel (ops.conv_dgrad), IS = 1, OS = 2, offset = parity
    dy2 = rs.randn(2, 4, oz, oy, ox)
    got = np.zeros((2, 3, Z, Y, X))
    for pz in range(2):
        for px in range(2):
            for py in range(2):
                rows = ((Z - pz + 1) // 2, (X - px + 1) // 2, (Y - py + 1) // 2)
                if min(rows) > 0:
                    got += emulate(dy2, dg, grid, ops.taps_dgrad_s2(pz, px, py), rows, IS=1, OS=2, off=(pz, px, py))
    ref = F.conv_transpose3d(torch.from_numpy(dy2), wt, stride=2, padding=1,
                             output_padding=(Z - (2 * oz - 1), Y - (2 * oy - 1), X - (2 * ox - 1)))
    assert ref.shape[2:] == grid
    np.testing.assert_allclose(got, ref.numpy(), atol=1e-12)


def test_pointwise_and_dcn_tap_lists():
    rs = np.random.RandomState(0)
    x = rs.randn(1, 5, 2, 3, 4)
    w = rs.randn(6, 5, 1, 1, 1)
    fwd, dg = packs(w)
    got = emulate(x, fwd, (2, 3, 4), ops.taps_fwd(1), (2, 4, 3))
    np.testing.assert_allclose(got, F.conv3d(torch.from_numpy(x), torch.from_numpy(w)).numpy(), atol=1e-12)
    # the DCN sample volume keeps its 9 taps on the z axis: taps {(t, 0, 0, t)} over rows (1, Wo, Ho) contract C*9 (dcn.py)
    S = rs.randn(1, 5, 9, 3, 4)                      # [N, C, taps, Ho, Wo]
    wd = rs.randn(6, 5, 3, 3)
    fwd9, _ = packs(wd.reshape(6, 5, 9, 1, 1))
    got = emulate(S, fwd9, (1, 3, 4), [(t, 0, 0, t) for t in range(9)], (1, 4, 3))
    ref = np.einsum("nctyx,oct->noyx", S, wd.reshape(6, 5, 9))
    np.testing.assert_allclose(got[:, :, 0], ref, atol=1e-12)


def emulate_wgrad(x, dy, taps, rows, IS=1):
    """rtp_wgrad: dW[tap][ci][co] = sum_rows X[row*IS + tap][ci] * dY[row][co] (zero outside X); rows index dY directly."""
    N, Ci, Z, Y, X = x.shape
    Co = dy.shape[1]
    dW = np.zeros((len(taps), Ci, Co))
    RZ, RX, RY = rows
    for i, (tz, tx, ty) in enumerate(t[:3] for t in taps):
        for rz in range(RZ):
            for rx in range(RX):
                for ry in range(RY):
                    z, xx, y = rz * IS + tz, rx * IS + tx, ry * IS + ty
                    if 0 <= z < Z and 0 <= xx < X and 0 <= y < Y:
                        dW[i] += x[:, :, z, y, xx].T @ dy[:, :, rz, ry, rx]
    return dW


@pytest.mark.parametrize("stride,grid", [(1, (4, 5, 6)), (2, (4, 6, 8)), (2, (5, 7, 9))])
def test_weight_gradient_descriptors(stride, grid):
    """ops.conv_wgrad: taps_fwd(3), rows = dY grid, IS = stride; the reduction kernel writes dW[co][ci][tap]."""
    rs = np.random.RandomState(stride + sum(grid))
    Z, Y, X = grid
    x = torch.from_numpy(rs.randn(2, 3, Z, Y, X))
    w = torch.from_numpy(rs.randn(4, 3, 3, 3, 3)).requires_grad_(True)
    out = F.conv3d(x, w, stride=stride, padding=1)
    dy = torch.from_numpy(rs.randn(*out.shape))
    out.backward(dy)
    oz, oy, ox = out.shape[2:]
    dW = emulate_wgrad(x.numpy(), dy.numpy(), ops.taps_fwd(3), (oz, ox, oy), IS=stride)
    np.testing.assert_allclose(dW.transpose(2, 1, 0).reshape(4, 3, 3, 3, 3), w.grad.numpy(), atol=1e-11)
