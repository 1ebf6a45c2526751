// fuse.cu — multi-resolution branch exchange: fused (sum of same-resolution terms + trilinear align_corners
// upsample of low-resolution terms + bias, ReLU) in one pass, its transpose for the backward pass, and the
// gradient pass-through of the sum.  HBM-bound; fp32 accumulation, one bf16 store per output vector.
//
// Interpolation follows ATen's upsample_trilinear3d (align_corners=True): scale = float(in-1)/float(out-1)
// (0 when out == 1), src = scale * dst, i0 = int(src), i1 = i0 + (i0 < in-1), w1 = src - i0, w0 = 1 - w1.
#include <cstdlib>

#include "common.cuh"

namespace {

struct Axis {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ __forceinline__ Axis ac_axis(int d, int in, float scale) {
  Axis a;
  const float src = scale * (float)d;
  a.i0 = (int)src;
  a.i1 = a.i0 + (a.i0 < in - 1 ? 1 : 0);
  a.w1 = src - (float)a.i0;
  a.w0 = 1.f - a.w1;
  return a;
}

struct FuseK {
  P8 out;
  int C, n_same, n_low, relu;
  P8 same[4];
  P8 low[3];
  const float* bias;
};

// Warp-per-row variant (all low-resolution rows <= 32 vectors): trilinear interpolation is separable, so the warp first
// blends the 4 (z, x) corner rows of every low-res term into ONE row held in shared memory (4*Yl loads) and then
// interpolates along y from there — 4*Yl instead of 8*Y global loads per term and output row.
__global__ void __launch_bounds__(256) fuse_sum_rows_kernel(const __grid_constant__ FuseK p) {
  __shared__ float srow[8][3][32][8];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const P8& o = p.out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float bias[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias[i] = (p.bias && c8 * 8 + i < p.C) ? p.bias[c8 * 8 + i] : 0.f;
  const int R = o.Z * o.X;
  const int r0 = (int)((int64_t)R * blockIdx.x / gridDim.x), r1 = (int)((int64_t)R * (blockIdx.x + 1) / gridDim.x);
  float sz[3], sx[3], sy[3];
  for (int j = 0; j < p.n_low; ++j) {
    sz[j] = ac_scale(p.low[j].Z, o.Z);
    sx[j] = ac_scale(p.low[j].X, o.X);
    sy[j] = ac_scale(p.low[j].Y, o.Y);
  }
  // With three low-resolution terms the rows of the two coarsest ones (Y1 + Y2 <= 32 vectors) are blended in ONE pass:
  // lanes [0, Y1) work on term 1, lanes [Y1, Y1+Y2) on term 2 (the blend costs ~225 instructions per pass whatever the
  // number of active lanes, and was 52 % of this kernel's instructions).
  const bool merge12 = p.n_low == 3 && p.low[1].Y + p.low[2].Y <= 32;
  const int Y1 = p.low[1].Y;
  const bool second = merge12 && lane >= Y1;
  const P8& lm = second ? p.low[2] : p.low[1];
  const int mj = second ? 2 : 1, myl = second ? lane - Y1 : lane;
  const float msz = second ? sz[2] : sz[1], msx = second ? sx[2] : sx[1];
  const bf16* mlb = lm.ptr + n * lm.n_stride + c8 * lm.c_stride;
  const int mZ = lm.Z, mX = lm.X, mXp = lm.Xp, mYp = lm.Yp;
  const bool mactive = merge12 && myl < lm.Y;
  for (int row = r0 + warp; row < r1; row += 8) {
    const int z = row / o.X, x = row - z * o.X;
    if (mactive) {
      const Axis az = ac_axis(z, mZ, msz), ax = ac_axis(x, mX, msx);
      auto vox = [&](int zz, int xx) { return (((int64_t)zz * mXp + (xx + 1)) * mYp + (myl + 1)) * 8; };
      float f00[8], f01[8], f10[8], f11[8];
      unpack8(ldg16(mlb + vox(az.i0, ax.i0)), f00);
      unpack8(ldg16(mlb + vox(az.i0, ax.i1)), f01);
      unpack8(ldg16(mlb + vox(az.i1, ax.i0)), f10);
      unpack8(ldg16(mlb + vox(az.i1, ax.i1)), f11);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        srow[warp][mj][myl][i] = az.w0 * (ax.w0 * f00[i] + ax.w1 * f01[i]) + az.w1 * (ax.w0 * f10[i] + ax.w1 * f11[i]);
    }
    for (int j = 0; j < (merge12 ? 1 : p.n_low); ++j) {
      const P8& l = p.low[j];
      const Axis az = ac_axis(z, l.Z, sz[j]), ax = ac_axis(x, l.X, sx[j]);
      const bf16* lb = l.ptr + n * l.n_stride + c8 * l.c_stride;
      for (int yl = lane; yl < l.Y; yl += 32) {
        float f00[8], f01[8], f10[8], f11[8];
        unpack8(ldg16(lb + l.voxel(az.i0, ax.i0, yl)), f00);
        unpack8(ldg16(lb + l.voxel(az.i0, ax.i1, yl)), f01);
        unpack8(ldg16(lb + l.voxel(az.i1, ax.i0, yl)), f10);
        unpack8(ldg16(lb + l.voxel(az.i1, ax.i1, yl)), f11);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          srow[warp][j][yl][i] = az.w0 * (ax.w0 * f00[i] + ax.w1 * f01[i]) + az.w1 * (ax.w0 * f10[i] + ax.w1 * f11[i]);
      }
    }
    __syncwarp();
    for (int y = lane; y < o.Y; y += 32) {
      const int64_t off = o.voxel(z, x, y);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int s = 0; s < p.n_same; ++s) {
        float f[8];
        unpack8(ldg16(p.same[s].ptr + n * p.same[s].n_stride + c8 * p.same[s].c_stride + off), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
      for (int j = 0; j < p.n_low; ++j) {
        const Axis ay = ac_axis(y, p.low[j].Y, sy[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += ay.w0 * srow[warp][j][ay.i0][i] + ay.w1 * srow[warp][j][ay.i1][i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += bias[i];
        if (p.relu) acc[i] = fmaxf(acc[i], 0.f);
      }
      stg16(o.ptr + n * o.n_stride + c8 * o.c_stride + off, pack8(acc));
    }
    __syncwarp();
  }
}

// Tile variant (output rows of <= 64 vectors, low-resolution rows of <= 32): a CTA owns XT consecutive x rows of one
// (n, chunk, z) plane.  Phase 1 blends, for every low-resolution term, the four (z, x) corner rows of each of the XT output
// rows into shared memory (fp32) — the z axis and the per-term pointers are CTA constants, so a blended vector costs 4 loads,
// 4 unpacks and 8 x 4 FMAs and nothing else.  Phase 2: thread = (row, y) with y fixed for the whole kernel, so the y-axis
// interpolation of every term (indices, weights) lives in registers; per output vector: the same-resolution loads, two
// shared-memory rows per term, bias / ReLU, one 16-byte store.  Same fp32 expressions as fuse_sum_rows_kernel (which spent
// ~3x the instructions on per-row address arithmetic, int<->float conversions and local-memory copies of the term table).
template <int NLOW>
__global__ void __launch_bounds__(256, 3) fuse_sum_tile_kernel(const __grid_constant__ FuseK p, int XT, int log2ty) {
  extern __shared__ float4 fs_smem[];  // per term j: [XT][Yl_j][2]
  const P8& o = p.out;
  const int z = blockIdx.y;
  const int c8 = blockIdx.z % ((p.C + 7) / 8), n = blockIdx.z / ((p.C + 7) / 8);
  const int x0 = blockIdx.x * XT, nx = min(XT, o.X - x0);
  const int tid = threadIdx.x;
  const bf16* lb[NLOW];
  int Yl[NLOW], Ypl[NLOW], Xl[NLOW], sbase[NLOW];
  int64_t zo0[NLOW], zo1[NLOW];
  float wz0[NLOW], wz1[NLOW], sx[NLOW];
  int sm = 0;
#pragma unroll
  for (int j = 0; j < NLOW; ++j) {
    const P8& l = p.low[j];
    lb[j] = l.ptr + n * l.n_stride + c8 * l.c_stride;
    Yl[j] = l.Y; Ypl[j] = l.Yp; Xl[j] = l.X;
    const Axis az = ac_axis(z, l.Z, ac_scale(l.Z, o.Z));
    zo0[j] = (int64_t)az.i0 * l.Xp * l.Yp * 8;
    zo1[j] = (int64_t)az.i1 * l.Xp * l.Yp * 8;
    wz0[j] = az.w0; wz1[j] = az.w1;
    sx[j] = ac_scale(l.X, o.X);
    sbase[j] = sm;
    sm += XT * l.Y * 2;
  }
  // the first same-resolution term of this thread's rows is requested BEFORE the blend phase, so its HBM latency is hidden
  // behind phase 1 and the barrier (a thread has at most kRows rows: XT <= 16, >= 4 rows of threads)
  constexpr int kRows = 4;
  const int y = tid & ((1 << log2ty) - 1), rstep = 256 >> log2ty, xr0 = tid >> log2ty;
  const bool yok = y < o.Y;
  const int64_t xstride = (int64_t)o.Yp * 8;
  const int64_t obase = n * o.n_stride + c8 * o.c_stride + o.voxel(z, x0, y);
  uint4 pre[kRows];
  {
    const bf16* sb = p.same[0].ptr + n * p.same[0].n_stride + c8 * p.same[0].c_stride + o.voxel(z, x0, y);
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int xr = xr0 + k * rstep;
      pre[k] = (p.n_same > 0 && yok && xr < nx) ? ldg16(sb + (int64_t)xr * xstride) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // ---- phase 1: corner blend of the low-resolution rows
#pragma unroll
  for (int j = 0; j < NLOW; ++j) {
    int l2 = 0;
    while ((1 << l2) < Yl[j]) ++l2;
    const int yl = tid & ((1 << l2) - 1), rs1 = 256 >> l2;
    if (yl < Yl[j]) {
      for (int xr = tid >> l2; xr < nx; xr += rs1) {
        const Axis ax = ac_axis(x0 + xr, Xl[j], sx[j]);
        const int64_t o0 = ((int64_t)(ax.i0 + 1) * Ypl[j] + (yl + 1)) * 8, o1 = ((int64_t)(ax.i1 + 1) * Ypl[j] + (yl + 1)) * 8;
        float f00[8], f01[8], f10[8], f11[8], r[8];
        unpack8(ldg16(lb[j] + zo0[j] + o0), f00);
        unpack8(ldg16(lb[j] + zo0[j] + o1), f01);
        unpack8(ldg16(lb[j] + zo1[j] + o0), f10);
        unpack8(ldg16(lb[j] + zo1[j] + o1), f11);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          r[i] = wz0[j] * (ax.w0 * f00[i] + ax.w1 * f01[i]) + wz1[j] * (ax.w0 * f10[i] + ax.w1 * f11[i]);
        float4* d = fs_smem + sbase[j] + (xr * Yl[j] + yl) * 2;
        d[0] = make_float4(r[0], r[1], r[2], r[3]);
        d[1] = make_float4(r[4], r[5], r[6], r[7]);
      }
    }
  }
  __syncthreads();
  // ---- phase 2
  if (!yok) return;
  int yi0[NLOW], yi1[NLOW];
  float wy0[NLOW], wy1[NLOW];
#pragma unroll
  for (int j = 0; j < NLOW; ++j) {
    const Axis ay = ac_axis(y, Yl[j], ac_scale(Yl[j], o.Y));
    yi0[j] = ay.i0 * 2; yi1[j] = ay.i1 * 2; wy0[j] = ay.w0; wy1[j] = ay.w1;
  }
  float bias[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias[i] = (p.bias && c8 * 8 + i < p.C) ? p.bias[c8 * 8 + i] : 0.f;
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int xr = xr0 + k * rstep;
    if (xr >= nx) break;
    const int64_t off = o.voxel(z, x0 + xr, y);
    float acc[8];
    unpack8(pre[k], acc);  // zeros when there is no same-resolution term
    for (int s = 1; s < p.n_same; ++s) {
      float f[8];
      unpack8(ldg16(p.same[s].ptr + n * p.same[s].n_stride + c8 * p.same[s].c_stride + off), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
#pragma unroll
    for (int j = 0; j < NLOW; ++j) {
      const float4* r = fs_smem + sbase[j] + xr * Yl[j] * 2;
      const float4 a0 = r[yi0[j]], a1 = r[yi0[j] + 1], b0 = r[yi1[j]], b1 = r[yi1[j] + 1];
      acc[0] += wy0[j] * a0.x + wy1[j] * b0.x; acc[1] += wy0[j] * a0.y + wy1[j] * b0.y;
      acc[2] += wy0[j] * a0.z + wy1[j] * b0.z; acc[3] += wy0[j] * a0.w + wy1[j] * b0.w;
      acc[4] += wy0[j] * a1.x + wy1[j] * b1.x; acc[5] += wy0[j] * a1.y + wy1[j] * b1.y;
      acc[6] += wy0[j] * a1.z + wy1[j] * b1.z; acc[7] += wy0[j] * a1.w + wy1[j] * b1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += bias[i];
      if (p.relu) acc[i] = fmaxf(acc[i], 0.f);
    }
    stg16(o.ptr + obase + (int64_t)xr * xstride, pack8(acc));
  }
}

__global__ void __launch_bounds__(256) fuse_sum_kernel(const __grid_constant__ FuseK p) {
  // rows (fixed z, x) x lanes along y: the z/x interpolation indices and weights are computed once per row
  const int c8 = blockIdx.y, n = blockIdx.z;
  const P8& o = p.out;
  float bias[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias[i] = (p.bias && c8 * 8 + i < p.C) ? p.bias[c8 * 8 + i] : 0.f;
  int log2ty = 3;
  while ((1 << log2ty) < o.Y && log2ty < 6) ++log2ty;
  const int TY = 1 << log2ty;
  const int ty = threadIdx.x & (TY - 1), tr = threadIdx.x >> log2ty, rstep = 256 >> log2ty;
  const int R = o.Z * o.X;
  const int r0 = (int)((int64_t)R * blockIdx.x / gridDim.x), r1 = (int)((int64_t)R * (blockIdx.x + 1) / gridDim.x);
  float sz[3], sx[3], sy[3];
  for (int j = 0; j < p.n_low; ++j) {
    sz[j] = ac_scale(p.low[j].Z, o.Z);
    sx[j] = ac_scale(p.low[j].X, o.X);
    sy[j] = ac_scale(p.low[j].Y, o.Y);
  }
  int row = r0 + tr;
  if (row >= r1) return;
  int z = row / o.X, x = row - z * o.X;
  for (; row < r1; row += rstep) {
    Axis az[3], ax[3];
    for (int j = 0; j < p.n_low; ++j) {
      az[j] = ac_axis(z, p.low[j].Z, sz[j]);
      ax[j] = ac_axis(x, p.low[j].X, sx[j]);
    }
    for (int y = ty; y < o.Y; y += TY) {
      const int64_t off = o.voxel(z, x, y);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int s = 0; s < p.n_same; ++s) {
        float f[8];
        unpack8(ldg16(p.same[s].ptr + n * p.same[s].n_stride + c8 * p.same[s].c_stride + off), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
      for (int j = 0; j < p.n_low; ++j) {
        const P8& l = p.low[j];
        const Axis ay = ac_axis(y, l.Y, sy[j]);
        const bf16* lb = l.ptr + n * l.n_stride + c8 * l.c_stride;
        float up[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) up[i] = 0.f;
#pragma unroll
        for (int corner = 0; corner < 8; ++corner) {
          const int zz = (corner & 4) ? az[j].i1 : az[j].i0;
          const int xx = (corner & 2) ? ax[j].i1 : ax[j].i0;
          const int yy = (corner & 1) ? ay.i1 : ay.i0;
          const float w = ((corner & 4) ? az[j].w1 : az[j].w0) * ((corner & 2) ? ax[j].w1 : ax[j].w0) * ((corner & 1) ? ay.w1 : ay.w0);
          float f[8];
          unpack8(ldg16(lb + l.voxel(zz, xx, yy)), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) up[i] = fmaf(w, f[i], up[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += up[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += bias[i];
        if (p.relu) acc[i] = fmaxf(acc[i], 0.f);
      }
      stg16(o.ptr + n * o.n_stride + c8 * o.c_stride + off, pack8(acc));
    }
    x += rstep;
    while (x >= o.X) { x -= o.X; ++z; }
  }
}

// Weight of source index l in the interpolation of destination d (ac_axis): a hat function of src = scale*d around l,
// written with the same fp32 expressions the forward uses (w0 = 1 - (src - i0), w1 = src - i0).
__device__ __forceinline__ float hat_weight(int d, int l, float scale) {
  const float src = scale * (float)d;
  const float w = src >= (float)l ? 1.f - (src - (float)l) : src - (float)(l - 1);
  return w > 0.f ? w : 0.f;
}

// Transpose of the trilinear upsample, applied separably: U^T = Uz^T Ux^T Uy^T.  One launch reduces ONE axis:
//   out[.., l, ..] (=|+=) sum_d w(d -> l) * in[.., d, ..]      (gather form, deterministic)
// so the full-resolution gradient is read ~once instead of once per low-res footprint (8x..27x).
// axis: 0 = z, 1 = x, 2 = y.  `in` and `out` differ only in the extent of `axis`.  The destinations touching l are
// the d with |scale*d - l| < 1; they are visited four at a time with clamped indices (weight 0 outside the range) so
// the loads of a group are independent and issue back to back.
// Launch geometry: grid (x tiles, z, n*C8 + c8), block = (2^log2ty lanes along y) x (256 >> log2ty rows along x) — no
// per-element index divisions.  The candidate range is exact: the conservative bounds from the reciprocal are tightened
// with the same `scale * d` expression the weights use, so a 2x pass reads one group of <= 4 vectors per output.
__global__ void __launch_bounds__(256) upsample_bwd_axis_kernel(P8 in, P8 out, int C8, int axis, int accumulate, int log2ty) {
  const int in_n = axis == 0 ? in.Z : (axis == 1 ? in.X : in.Y);
  const int out_n = axis == 0 ? out.Z : (axis == 1 ? out.X : out.Y);
  const float scale = ac_scale(out_n, in_n);  // low-res (out) is the interpolation source, high-res (in) the dest
  const float inv = scale > 0.f ? 1.f / scale : 0.f;
  const int64_t astride = axis == 0 ? in.plane_elems() : (axis == 1 ? (int64_t)in.Yp * 8 : 8);
  const int TY = 1 << log2ty, TX = 256 >> log2ty;
  const int ty = threadIdx.x & (TY - 1), tx = threadIdx.x >> log2ty;
  const int z = blockIdx.y;
  const int c8 = blockIdx.z % C8, n = blockIdx.z / C8;
  const int x = blockIdx.x * TX + tx;
  if (x >= out.X) return;
  const bf16* in_nc = in.ptr + n * in.n_stride + c8 * in.c_stride;
  bf16* out_nc = out.ptr + n * out.n_stride + c8 * out.c_stride;
  for (int y = ty; y < out.Y; y += TY) {
    const int l = axis == 0 ? z : (axis == 1 ? x : y);
    int lo = 0, hi = in_n - 1;
    if (scale > 0.f) {
      lo = max(0, (int)floorf((float)(l - 1) * inv));
      hi = min(in_n - 1, (int)ceilf((float)(l + 1) * inv));
      if (scale * (float)lo <= (float)(l - 1)) ++lo;  // weight exactly 0 there
      if (scale * (float)hi >= (float)(l + 1)) --hi;
    }
    // voxel of this output with the reduced axis at 0
    const bf16* ib = in_nc + in.voxel(axis == 0 ? 0 : z, axis == 1 ? 0 : x, axis == 2 ? 0 : y);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int d0 = lo; d0 <= hi; d0 += 4) {
      uint4 v[4];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = min(d0 + u, hi);
        v[u] = ldg16(ib + d * astride);
        w[u] = d0 + u <= hi ? (scale > 0.f ? hat_weight(d, l, scale) : 1.f) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[u], f[k], acc[k]);
      }
    }
    bf16* dst = out_nc + out.voxel(z, x, y);
    if (accumulate) {
      float g[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += g[k];
    }
    stg16(dst, pack8(acc));
  }
}

// Exact candidate range of destinations d with non-zero weight for source index l (see upsample_bwd_axis_kernel).
__device__ __forceinline__ void hat_range(int l, int in_n, float scale, float inv, int& lo, int& hi) {
  lo = max(0, (int)floorf((float)(l - 1) * inv));
  hi = min(in_n - 1, (int)ceilf((float)(l + 1) * inv));
  if (scale * (float)lo <= (float)(l - 1)) ++lo;
  if (scale * (float)hi >= (float)(l + 1)) --hi;
}

// The y and x reductions of U^T in ONE launch: a CTA owns kTXL low-resolution x positions of one (n, chunk, z) plane.
// Phase 1 reduces every full-resolution row it needs along y into shared memory (fp32), phase 2 reduces those rows along
// x.  The [Z][X][Yl] intermediate of the separable scheme never goes to HBM (a third of its traffic at ratio 2) and is not
// rounded to bf16 in between.  Requires both scales > 0 (low extents > 1); the host falls back to two axis launches otherwise.
constexpr int kTXL = 8;
__global__ void __launch_bounds__(256) upsample_bwd_yx_kernel(P8 in, P8 out, int C8, int nrows_max) {
  extern __shared__ float4 up_smem[];  // [nrows_max][Yl][2] float4
  const float sy = ac_scale(out.Y, in.Y), sx = ac_scale(out.X, in.X);
  const float iy = 1.f / sy, ix = 1.f / sx;
  const int z = blockIdx.y, c8 = blockIdx.z % C8, n = blockIdx.z / C8;
  const int xl0 = blockIdx.x * kTXL, xl1 = min(out.X, xl0 + kTXL) - 1;
  int r0, r1, tmp;
  hat_range(xl0, in.X, sx, ix, r0, tmp);
  hat_range(xl1, in.X, sx, ix, tmp, r1);
  const int nrows = min(r1 - r0 + 1, nrows_max);
  const int Yl = out.Y;
  const bf16* in_nc = in.ptr + n * in.n_stride + c8 * in.c_stride;
  for (int i = threadIdx.x; i < nrows * Yl; i += blockDim.x) {
    const int r = i / Yl, yl = i - r * Yl;
    int lo, hi;
    hat_range(yl, in.Y, sy, iy, lo, hi);
    const bf16* ib = in_nc + in.voxel(z, r0 + r, 0);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int d0 = lo; d0 <= hi; d0 += 4) {
      uint4 v[4];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = min(d0 + u, hi);
        v[u] = ldg16(ib + d * 8);
        w[u] = d0 + u <= hi ? hat_weight(d, yl, sy) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[u], f[k], acc[k]);
      }
    }
    up_smem[2 * i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    up_smem[2 * i + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  bf16* out_nc = out.ptr + n * out.n_stride + c8 * out.c_stride;
  const int nxl = xl1 - xl0 + 1;
  for (int i = threadIdx.x; i < nxl * Yl; i += blockDim.x) {
    const int xi = i / Yl, yl = i - xi * Yl, xl = xl0 + xi;
    int lo, hi;
    hat_range(xl, in.X, sx, ix, lo, hi);
    hi = min(hi, r0 + nrows - 1);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int d = lo; d <= hi; ++d) {
      const float w = hat_weight(d, xl, sx);
      const float4 a = up_smem[2 * ((d - r0) * Yl + yl)], b = up_smem[2 * ((d - r0) * Yl + yl) + 1];
      acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
      acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    }
    stg16(out_nc + out.voxel(z, xl, yl), pack8(acc));
  }
}

// Warp-cooperative variant of upsample_bwd_yx_kernel for Y == B * Yl (B = 2, 4, 8; Yl a power of two <= 32).
// ncu of the gather version (profiles/r02_ncu_elementwise.txt): 80 % L1/TEX throughput, 31 % DRAM — consecutive lanes
// (consecutive yl) gathered full-resolution vectors B apart, and every vector was requested by two or three hats: ~4 L1
// wavefronts per input vector.  Here lane l of a row group loads the B consecutive vectors [l*B, (l+1)*B) of the row — a
// fully coalesced read, each vector requested ONCE — and the hat of yl, whose support lies inside the blocks of lanes
// l-1, l, l+1 (tests/test_upsample_math.py), is evaluated from its own registers and warp shuffles.  The hat weights depend
// only on (l, offset), so they are computed once per thread.  Candidates are visited in ascending d with the same fma chain as
// the gather kernel (zero-weight candidates leave the sum unchanged), so the result is bit-identical.
template <int B>
__global__ void __launch_bounds__(256) upsample_bwd_yx_shfl_kernel(P8 in, P8 out, int C8, int nrows_max) {
  extern __shared__ float4 up_smem[];  // [nrows_max][Yl][2] float4
  const float sy = ac_scale(out.Y, in.Y), sx = ac_scale(out.X, in.X);
  const float ix = 1.f / sx;
  const int z = blockIdx.y, c8 = blockIdx.z % C8, n = blockIdx.z / C8;
  const int xl0 = blockIdx.x * kTXL, xl1 = min(out.X, xl0 + kTXL) - 1;
  int r0, r1, tmp;
  hat_range(xl0, in.X, sx, ix, r0, tmp);
  hat_range(xl1, in.X, sx, ix, tmp, r1);
  const int nrows = min(r1 - r0 + 1, nrows_max);
  const int Yl = out.Y, Y = in.Y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / Yl, g = lane / Yl, l = lane - g * Yl;  // rows per warp, row group and yl of this lane
  float wgt[3][B];
  int srcl[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    const int ln = l + o - 1;
    srcl[o] = g * Yl + min(max(ln, 0), Yl - 1);
#pragma unroll
    for (int k = 0; k < B; ++k) {
      const int d = ln * B + k;
      wgt[o][k] = (ln >= 0 && ln < Yl && d < Y) ? hat_weight(d, l, sy) : 0.f;
    }
  }
  const bf16* in_nc = in.ptr + n * in.n_stride + c8 * in.c_stride + in.voxel(z, r0, 0);
  const int64_t rstride = (int64_t)in.Yp * 8;
  for (int rb = warp * rpw; rb < nrows; rb += 8 * rpw) {
    const int r = rb + g;
    const bool valid = r < nrows;
    uint4 raw[B];
#pragma unroll
    for (int k = 0; k < B; ++k) raw[k] = valid ? ldg16(in_nc + r * rstride + (l * B + k) * 8) : make_uint4(0u, 0u, 0u, 0u);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
#pragma unroll
      for (int k = 0; k < B; ++k) {
        uint4 v = raw[k];
        if (o != 1) {
          v.x = __shfl_sync(0xffffffffu, raw[k].x, srcl[o]);
          v.y = __shfl_sync(0xffffffffu, raw[k].y, srcl[o]);
          v.z = __shfl_sync(0xffffffffu, raw[k].z, srcl[o]);
          v.w = __shfl_sync(0xffffffffu, raw[k].w, srcl[o]);
        }
        float f[8];
        unpack8(v, f);
        const float w = wgt[o][k];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(w, f[q], acc[q]);
      }
    }
    if (valid) {
      up_smem[2 * (r * Yl + l)] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      up_smem[2 * (r * Yl + l) + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  __syncthreads();
  bf16* out_nc = out.ptr + n * out.n_stride + c8 * out.c_stride;
  const int nxl = xl1 - xl0 + 1;
  for (int i = threadIdx.x; i < nxl * Yl; i += blockDim.x) {
    const int xi = i / Yl, yl = i - xi * Yl, xl = xl0 + xi;
    int lo, hi;
    hat_range(xl, in.X, sx, ix, lo, hi);
    hi = min(hi, r0 + nrows - 1);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int d = lo; d <= hi; ++d) {
      const float w = hat_weight(d, xl, sx);
      const float4 a = up_smem[2 * ((d - r0) * Yl + yl)], b = up_smem[2 * ((d - r0) * Yl + yl) + 1];
      acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
      acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    }
    stg16(out_nc + out.voxel(z, xl, yl), pack8(acc));
  }
}

// dst (=|+=) src [* (mask > 0)]
__global__ void __launch_bounds__(256) grad_add_kernel(P8 src, P8 mask, int has_mask, P8 dst, int accumulate) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t V = (int64_t)dst.Z * dst.X * dst.Y;
  for (int64_t v = blockIdx.x * 256ll + threadIdx.x; v < V; v += (int64_t)gridDim.x * 256) {
    uint32_t q = (uint32_t)v;
    const int y = (int)(q % (uint32_t)dst.Y); q /= (uint32_t)dst.Y;
    const int x = (int)(q % (uint32_t)dst.X);
    const int z = (int)(q / (uint32_t)dst.X);
    const int64_t off = dst.voxel(z, x, y);
    float f[8];
    unpack8(ldg16(src.ptr + n * src.n_stride + c8 * src.c_stride + off), f);
    if (has_mask) {
      float m[8];
      unpack8(ldg16(mask.ptr + n * mask.n_stride + c8 * mask.c_stride + off), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = m[i] > 0.f ? f[i] : 0.f;
    }
    bf16* d = dst.ptr + n * dst.n_stride + c8 * dst.c_stride + off;
    if (accumulate) {
      float g[8];
      unpack8(*reinterpret_cast<const uint4*>(d), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
    stg16(d, pack8(f));
  }
}

int ew_blocks(int64_t V) {  // every thread gets >= ~8 vectors
  int64_t b = (V + 2047) / 2048;
  return (int)(b > 592 ? 592 : (b < 1 ? 1 : b));
}
bool same_geom(const rtp_p8& a, const rtp_p8& b) { return a.N == b.N && a.Z == b.Z && a.X == b.X && a.Y == b.Y; }

}  // namespace

int rtp_fuse_sum_mma(const rtp_fuse_desc* d, void* stream);  // fuse_mma.cu

extern "C" int rtp_fuse_sum(const rtp_fuse_desc* d, void* stream) {
  RTP_CHECK_ARG(d && d->out.ptr, "rtp_fuse_sum: null argument");
  RTP_CHECK_ARG(d->n_same >= 0 && d->n_same <= 4 && d->n_low >= 0 && d->n_low <= 3 && d->n_same + d->n_low >= 1,
                "rtp_fuse_sum: bad term counts");
  const int C8 = ceil_div(d->C, 8);
  RTP_CHECK_ARG(d->C > 0 && C8 <= d->out.C8, "rtp_fuse_sum: bad C");
  FuseK k;
  k.out = P8(d->out); k.C = d->C; k.n_same = d->n_same; k.n_low = d->n_low; k.relu = d->relu; k.bias = d->bias;
  for (int i = 0; i < d->n_same; ++i) {
    RTP_CHECK_ARG(d->same[i].ptr && same_geom(d->same[i], d->out) && d->same[i].C8 >= C8, "rtp_fuse_sum: same[%d] geometry mismatch", i);
    k.same[i] = P8(d->same[i]);
  }
  for (int i = 0; i < d->n_low; ++i) {
    RTP_CHECK_ARG(d->low[i].ptr && d->low[i].N == d->out.N && d->low[i].C8 >= C8, "rtp_fuse_sum: low[%d] mismatch", i);
    k.low[i] = P8(d->low[i]);
  }
  {
    // y interpolation on the tensor cores (fuse_mma.cu); 0 = shape not supported
    const int r = rtp_fuse_sum_mma(d, stream);
    if (r < 0) return -1;
    if (r > 0) RTP_LAUNCH_CHECK();
  }
  const int64_t V = (int64_t)d->out.Z * d->out.X * d->out.Y;
  bool rows_ok = d->n_low > 0;
  for (int i = 0; i < d->n_low; ++i) rows_ok = rows_ok && d->low[i].Y <= 32;
  static const bool no_tile = getenv("RTP_NO_FUSE_TILE") != nullptr;  // A/B switch
  if (rows_ok && d->out.Y <= 64 && !no_tile) {
    int sumY = 0;
    for (int i = 0; i < d->n_low; ++i) sumY += d->low[i].Y;
    const int XT = 16 * sumY * 32 <= 32 * 1024 ? 16 : 8;
    const size_t smem = (size_t)XT * sumY * 32;
    int log2ty = 0;
    while ((1 << log2ty) < d->out.Y) ++log2ty;  // <= 6: at least 4 rows of threads, so a thread has <= XT / 4 <= 4 rows
    const dim3 grid((unsigned)ceil_div(d->out.X, XT), (unsigned)d->out.Z, (unsigned)(d->out.N * C8));
    if (d->n_low == 1) fuse_sum_tile_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(k, XT, log2ty);
    else if (d->n_low == 2) fuse_sum_tile_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(k, XT, log2ty);
    else fuse_sum_tile_kernel<3><<<grid, 256, smem, (cudaStream_t)stream>>>(k, XT, log2ty);
  } else if (rows_ok)
    fuse_sum_rows_kernel<<<dim3(ew_blocks(V), C8, d->out.N), 256, 0, (cudaStream_t)stream>>>(k);
  else
    fuse_sum_kernel<<<dim3(ew_blocks(V), C8, d->out.N), 256, 0, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

int rtp_upsample_bwd_yx_mma(const rtp_p8& dout, const rtp_p8& t2, int C8, void* stream);  // upsample_mma.cu

extern "C" int64_t rtp_upsample_bwd_workspace_bytes(rtp_p8 dout, rtp_p8 dlow, int32_t C) {
  // two intermediates: [Z][X][Yl] and [Z][Xl][Yl] (P8, padded planes)
  const int64_t C8 = ceil_div(C, 8);
  const int64_t a = (int64_t)dout.N * C8 * dout.Z * (dout.X + 2) * (dlow.Y + 2) * 16;
  const int64_t b = (int64_t)dout.N * C8 * dout.Z * (dlow.X + 2) * (dlow.Y + 2) * 16;
  return a + b + 512;
}

extern "C" int rtp_upsample_bwd(rtp_p8 dout, rtp_p8 dlow, int32_t C, int32_t accumulate, void* workspace, void* stream) {
  RTP_CHECK_ARG(dout.ptr && dlow.ptr && workspace && dout.N == dlow.N, "rtp_upsample_bwd: bad tensors");
  const int C8 = ceil_div(C, 8);
  RTP_CHECK_ARG(C > 0 && C8 <= dout.C8 && C8 <= dlow.C8, "rtp_upsample_bwd: bad C");
  RTP_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "rtp_upsample_bwd: workspace must be 16-byte aligned");
  // intermediates are dense P8 tensors carved from the workspace (their pad rings are never read)
  rtp_p8 t1 = dout, t2 = dout;
  t1.ptr = workspace; t1.C8 = C8; t1.Y = dlow.Y;
  t1.c_stride = (int64_t)t1.Z * (t1.X + 2) * (t1.Y + 2) * 8; t1.n_stride = t1.c_stride * C8;
  t2.ptr = (char*)workspace + (size_t)t1.N * t1.n_stride * 2; t2.C8 = C8; t2.Y = dlow.Y; t2.X = dlow.X;
  t2.c_stride = (int64_t)t2.Z * (t2.X + 2) * (t2.Y + 2) * 8; t2.n_stride = t2.c_stride * C8;
  auto launch = [&](const rtp_p8& a, const rtp_p8& b, int axis, int acc) {
    int log2ty = 2;
    while ((1 << log2ty) < b.Y && log2ty < 6) ++log2ty;
    const int TX = 256 >> log2ty;
    upsample_bwd_axis_kernel<<<dim3((unsigned)ceil_div(b.X, TX), (unsigned)b.Z, (unsigned)(b.N * C8)), 256, 0, (cudaStream_t)stream>>>(
        P8(a), P8(b), C8, axis, acc, log2ty);
  };
  static const bool no_fused = getenv("RTP_NO_FUSED_UPBWD") != nullptr;  // A/B switch
  const float sxh = dlow.X > 1 && dout.X > 1 ? (float)(dlow.X - 1) / (float)(dout.X - 1) : 0.f;
  // rows a tile can need: hi(l1) - lo(l0) + 1 <= (kTXL + 1) / scale + 3 in exact arithmetic; one more for fp32 rounding
  const int nrows_max = sxh > 0.f ? (int)ceilf((float)(kTXL + 1) / sxh) + 4 : 0;
  const size_t smem = (size_t)nrows_max * dlow.Y * 32;
  int mma_taken = 0;
  if (!no_fused && sxh > 0.f && dlow.Y > 1 && dout.Y > 1) {
    // y reduction on the tensor cores, x reduction thread-local (upsample_mma.cu); 0 = shape not supported
    mma_taken = rtp_upsample_bwd_yx_mma(dout, t2, C8, stream);
    if (mma_taken < 0) return -1;
  }
  if (mma_taken) {
  } else if (!no_fused && sxh > 0.f && dlow.Y > 1 && dout.Y > 1 && smem <= 96 * 1024) {
    static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
    if (smem > configured) {
      cudaFuncSetAttribute(upsample_bwd_yx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = smem;
    }
    static const bool no_shfl = getenv("RTP_NO_UPBWD_SHFL") != nullptr;  // A/B switch
    const int Bq = dlow.Y > 0 ? dout.Y / dlow.Y : 0;
    const bool pow2 = (dlow.Y & (dlow.Y - 1)) == 0;
    const dim3 grid((unsigned)ceil_div(t2.X, kTXL), (unsigned)t2.Z, (unsigned)(t2.N * C8));
    if (!no_shfl && smem <= 48 * 1024 && pow2 && dlow.Y <= 32 && dlow.Y >= 4 && Bq * dlow.Y == dout.Y && (Bq == 2 || Bq == 4)) {
      // (B = 8 — 24 candidate vectors per output, 103 registers — measured slower than the gather kernel: 0.103 vs 0.083 ms)
      if (Bq == 2) upsample_bwd_yx_shfl_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(P8(dout), P8(t2), C8, nrows_max);
      else upsample_bwd_yx_shfl_kernel<4><<<grid, 256, smem, (cudaStream_t)stream>>>(P8(dout), P8(t2), C8, nrows_max);
    } else {
      upsample_bwd_yx_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(P8(dout), P8(t2), C8, nrows_max);
    }
  } else {
    launch(dout, t1, 2, 0);
    launch(t1, t2, 1, 0);
  }
  launch(t2, dlow, 0, accumulate);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_grad_add(rtp_p8 src, rtp_p8 mask, rtp_p8 dst, int32_t C, int32_t accumulate, void* stream) {
  RTP_CHECK_ARG(src.ptr && dst.ptr && same_geom(src, dst), "rtp_grad_add: geometry mismatch");
  const int C8 = ceil_div(C, 8);
  RTP_CHECK_ARG(C > 0 && C8 <= src.C8 && C8 <= dst.C8, "rtp_grad_add: bad C");
  if (mask.ptr) RTP_CHECK_ARG(same_geom(mask, dst) && C8 <= mask.C8, "rtp_grad_add: mask geometry mismatch");
  const int64_t V = (int64_t)dst.Z * dst.X * dst.Y;
  grad_add_kernel<<<dim3(ew_blocks(V), C8, dst.N), 256, 0, (cudaStream_t)stream>>>(P8(src), P8(mask), mask.ptr != nullptr,
                                                                                    P8(dst), accumulate);
  RTP_LAUNCH_CHECK();
}
