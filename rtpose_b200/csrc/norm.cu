// norm.cu — GroupNorm(G, C) forward / backward on P8 tensors.  HBM-bound elementwise + two-level
// (deterministic) reductions; statistics fp32, combined in fp64.
//
// Algorithmic bytes (bf16): sums 2 B/elem read; apply 2 B read + 2 B write; bwd_reduce 4 B read;
// bwd_apply 4 B read + 2 B write (+2 B when accumulating).
#include "common.cuh"

namespace {

constexpr int kSlabs = 32;  // partial reductions per (sample, chunk); fixed => deterministic summation order

// decode a linear real-voxel index v in [0, Z*X*Y) -> element offset (y fastest)
__device__ __forceinline__ int64_t voxel_off(const P8& t, int64_t v) {
  // 32-bit arithmetic on purpose (Z*X*Y < 2^31): 64-bit div/mod costs ~10x more and dominated these kernels
  uint32_t u = (uint32_t)v;
  const uint32_t y = u % (uint32_t)t.Y;
  u /= (uint32_t)t.Y;
  const uint32_t x = u % (uint32_t)t.X;
  const uint32_t z = u / (uint32_t)t.X;
  return t.voxel((int)z, (int)x, (int)y);
}

// Row-wise traversal of the real voxels: a row is Y consecutive 16-byte vectors.  Threads are laid out as
// (TY = 2^log2ty lanes along y) x (blockDim/TY rows); (z, x) advance incrementally, so there is ONE integer
// division per thread instead of two per vector.  f(element_offset) is called for every assigned vector.
template <typename F>
__device__ __forceinline__ void rows_foreach(const P8& t, int row_begin, int row_end, int log2ty, F&& f) {
  const int TY = 1 << log2ty;
  const int ty = threadIdx.x & (TY - 1), tr = threadIdx.x >> log2ty;
  const int rstep = (int)blockDim.x >> log2ty;
  int row = row_begin + tr;
  if (row < row_end) {
    int z = row / t.X, x = row - z * t.X;
    for (; row < row_end; row += rstep) {
      const int64_t base = t.voxel(z, x, 0);
      for (int y = ty; y < t.Y; y += TY) f(base + (int64_t)y * 8);
      x += rstep;
      while (x >= t.X) { x -= t.X; ++z; }
    }
  }
}
// same walk, also handing out the voxel coordinates: f(element_offset, z, x, y)
template <typename F>
__device__ __forceinline__ void rows_foreach_zxy(const P8& t, int row_begin, int row_end, int log2ty, F&& f) {
  const int TY = 1 << log2ty;
  const int ty = threadIdx.x & (TY - 1), tr = threadIdx.x >> log2ty;
  const int rstep = (int)blockDim.x >> log2ty;
  int row = row_begin + tr;
  if (row < row_end) {
    int z = row / t.X, x = row - z * t.X;
    for (; row < row_end; row += rstep) {
      const int64_t base = t.voxel(z, x, 0);
      for (int y = ty; y < t.Y; y += TY) f(base + (int64_t)y * 8, z, x, y);
      x += rstep;
      while (x >= t.X) { x -= t.X; ++z; }
    }
  }
}
// Space-to-depth view of a tensor with C8 chunks: voxel (z, x, y), chunk c8 lives in the half-resolution tensor `s` at
// voxel (z/2, x/2, y/2), chunk ((z&1)*4 + (x&1)*2 + (y&1)) * C8 + c8.  A stride-2 3x3x3 conv over the original is a
// stride-1 conv over this view (tap k = 0 / 1 / 2 of a dimension reads parity 1 at offset -1 / parity 0 at 0 / parity 1
// at 0), which is what lets the stride-2 exchange convs run on the plane-streaming kernels.
__device__ __forceinline__ int64_t s2d_offset(const P8& s, int C8, int c8, int z, int x, int y) {
  const int par = ((z & 1) << 2) | ((x & 1) << 1) | (y & 1);
  return (int64_t)(par * C8 + c8) * s.c_stride + s.voxel(z >> 1, x >> 1, y >> 1);
}
// Same walk with the loads of U rows issued before any of them is consumed: `ld(offset, z, x, y)` returns the loaded vectors,
// `use(offset, data)` computes / stores.  Without this the compiler cannot hoist the next row's loads above the current
// row's store (possible aliasing), and each thread has a single 16-byte load pair in flight.
// PAIR: a lane handles the two voxels (y, y+1), y even (t.Y even): used by the space-to-depth variants, where the two
// voxels live in different parity groups of the view and each group is then accessed with unit stride by the warp.
template <int U, bool PAIR = false, typename L, typename F>
__device__ __forceinline__ void rows_foreach_u(const P8& t, int row_begin, int row_end, int log2ty, L&& ld, F&& use) {
  const int TY = 1 << log2ty;
  const int ty = threadIdx.x & (TY - 1), tr = threadIdx.x >> log2ty;
  const int rstep = (int)blockDim.x >> log2ty;
  using D = decltype(ld((int64_t)0, 0, 0, 0));
  for (int y = PAIR ? 2 * ty : ty; y < t.Y; y += PAIR ? 2 * TY : TY) {
    int row = row_begin + tr;
    for (; row + (U - 1) * rstep < row_end; row += U * rstep) {
      int64_t off[U];
      D d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = row + u * rstep;
        const int z = r / t.X, x = r - z * t.X;
        off[u] = t.voxel(z, x, y);
        d[u] = ld(off[u], z, x, y);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) use(off[u], d[u]);
    }
    for (; row < row_end; row += rstep) {
      const int z = row / t.X, x = row - z * t.X;
      const int64_t off = t.voxel(z, x, y);
      use(off, ld(off, z, x, y));
    }
  }
}
struct Vec2 {
  uint4 a, b;
};
struct Vec3 {
  uint4 a, b, c;
};
struct Vec4 {
  uint4 a, b, c, d;
};
struct Vec6 {
  uint4 a, b, c, d, e, f;
};
struct Vec8 {
  uint4 a, b, c, d, e, f, g, h;
};
__host__ __device__ inline int log2_ty(int Y) {
  int l = 3;
  while ((1 << l) < Y && l < 6) ++l;
  return l;
}

// block-wide sum of NV values per thread; result valid in thread 0..NV-1 of warp 0 (value index = lane)
template <int NV>
__device__ __forceinline__ void block_reduce(float (&acc)[NV], float* sh /* [8][NV] */, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[warp * NV + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w * NV + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// partial[n][c8][slab][16] = (sum x[0..7], sum x^2[0..7])
__global__ void __launch_bounds__(256) gn_sums_partial_kernel(P8 x, float* __restrict__ partial) {
  __shared__ float sh[8 * 16];
  const int slab = blockIdx.x, c8 = blockIdx.y, n = blockIdx.z;
  const int R = x.Z * x.X;
  const int r0 = (int)((int64_t)R * slab / kSlabs), r1 = (int)((int64_t)R * (slab + 1) / kSlabs);
  const bf16* base = x.ptr + n * x.n_stride + c8 * x.c_stride;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  rows_foreach_u<4>(
      x, r0, r1, log2_ty(x.Y), [&](int64_t off, int, int, int) { return ldg16(base + off); },
      [&](int64_t, const uint4& v) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] += f[i];
          acc[8 + i] += f[i] * f[i];
        }
      });
  block_reduce<16>(acc, sh, partial + (((size_t)n * x.C8 + c8) * kSlabs + slab) * 16);
}

// sums[n][c][2] = sum over slabs (fixed order)
__global__ void gn_sums_final_kernel(const float* __restrict__ partial, int N, int C8, int C, float* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C8 * 8) return;
  const int j = i & 7, c8 = (i >> 3) % C8, n = (i >> 3) / C8;
  const int c = c8 * 8 + j;
  if (c >= C) return;
  const float* p = partial + ((size_t)n * C8 + c8) * kSlabs * 16;
  double s = 0, q = 0;
  for (int k = 0; k < kSlabs; ++k) {
    s += p[k * 16 + j];
    q += p[k * 16 + 8 + j];
  }
  sums[((size_t)n * C + c) * 2] = (float)s;
  sums[((size_t)n * C + c) * 2 + 1] = (float)q;
}

// stats[n][g] = (mean, rstd) from per-channel sums
__global__ void gn_finalize_kernel(const float* __restrict__ sums, int N, int C, int G, double count, float eps,
                                   float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * G) return;
  const int g = i % G, n = i / G, cpg = C / G;
  double s = 0, q = 0;
  for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
    s += sums[((size_t)n * C + c) * 2];
    q += sums[((size_t)n * C + c) * 2 + 1];
  }
  const double m = count * cpg;
  const double mean = s / m;
  double var = q / m - mean * mean;
  if (var < 0) var = 0;
  stats[i * 2] = (float)mean;
  stats[i * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// y = (x - mean) * rstd * gamma + beta  on real voxels (pads are never written)
template <bool S2D>  // S2D: y is the space-to-depth view (half resolution, 8x the chunks)
__global__ void __launch_bounds__(256) gn_apply_kernel(P8 x, int C, int G, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       P8 y) {
  __shared__ float s_ab[16];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int cpg = C / G;
  if (threadIdx.x < 8) {  // per-(sample, chunk) affine coefficients, computed once per block
    const int c = c8 * 8 + threadIdx.x;
    float av = 0.f, bv = 0.f;
    if (c < C) {
      const int g = c / cpg;
      const float mean = stats[((size_t)n * G + g) * 2], rstd = stats[((size_t)n * G + g) * 2 + 1];
      av = rstd * gamma[c];
      bv = beta[c] - mean * av;
    }
    s_ab[threadIdx.x] = av;
    s_ab[8 + threadIdx.x] = bv;
  }
  __syncthreads();
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = s_ab[i]; b[i] = s_ab[8 + i]; }
  const int R = x.Z * x.X;
  const int r0 = (int)((int64_t)R * blockIdx.x / gridDim.x), r1 = (int)((int64_t)R * (blockIdx.x + 1) / gridDim.x);
  const bf16* xb = x.ptr + n * x.n_stride + c8 * x.c_stride;
  if constexpr (S2D) {
    bf16* yb = y.ptr + n * y.n_stride;
    const int C8 = (int)gridDim.y;
    const int64_t pstride = (int64_t)C8 * y.c_stride;  // y -> y+1 (y even): next parity group of the view
    rows_foreach_u<2, true>(
        x, r0, r1, log2_ty(x.Y / 2),
        [&](int64_t off, int z, int xx, int yy) {
          Vec3 v;  // .c carries the destination offset inside the view
          v.a = ldg16(xb + off);
          v.b = ldg16(xb + off + 8);
          const int64_t so = s2d_offset(y, C8, c8, z, xx, yy);
          v.c = make_uint4((uint32_t)so, (uint32_t)(so >> 32), 0u, 0u);
          return v;
        },
        [&](int64_t, const Vec3& v) {
          float f[8], g[8];
          unpack8(v.a, f);
          unpack8(v.b, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[i] = fmaf(f[i], a[i], b[i]);
            g[i] = fmaf(g[i], a[i], b[i]);
          }
          const int64_t so = (int64_t)(((uint64_t)v.c.y << 32) | v.c.x);
          stg16(yb + so, pack8(f));
          stg16(yb + so + pstride, pack8(g));
        });
  } else {
    bf16* yb = y.ptr + n * y.n_stride + c8 * y.c_stride;
    rows_foreach_u<4>(
        x, r0, r1, log2_ty(x.Y), [&](int64_t off, int, int, int) { return ldg16(xb + off); },
        [&](int64_t off, const uint4& v) {
          float f[8];
          unpack8(v, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
          stg16(yb + off, pack8(f));
        });
  }
}

// partial[n][c8][slab][16] = (sum dy[0..7], sum dy*xhat[0..7])
template <bool S2D>  // S2D: dy is laid out as the space-to-depth view of x's grid
__global__ void __launch_bounds__(256) gn_bwd_partial_kernel(P8 x, P8 dy, int C, int G, const float* __restrict__ stats,
                                                             float* __restrict__ partial) {
  __shared__ float sh[8 * 16];
  const int slab = blockIdx.x, c8 = blockIdx.y, n = blockIdx.z;
  const int cpg = C / G;
  float mean[8], rstd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = min(c8 * 8 + i, C - 1);
    const int g = c / cpg;
    mean[i] = stats[((size_t)n * G + g) * 2];
    rstd[i] = stats[((size_t)n * G + g) * 2 + 1];
  }
  const int R = x.Z * x.X;
  const int r0 = (int)((int64_t)R * slab / kSlabs), r1 = (int)((int64_t)R * (slab + 1) / kSlabs);
  const bf16* xb = x.ptr + n * x.n_stride + c8 * x.c_stride;
  const bf16* db = dy.ptr + n * dy.n_stride + (S2D ? 0 : c8 * dy.c_stride);
  const int C8 = (int)gridDim.y;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  auto accum = [&](const uint4& xv, const uint4& dv) {
    float f[8], d[8];
    unpack8(xv, f);
    unpack8(dv, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += d[i];
      acc[8 + i] += d[i] * ((f[i] - mean[i]) * rstd[i]);
    }
  };
  if constexpr (S2D) {
    const int64_t pstride = (int64_t)C8 * dy.c_stride;
    rows_foreach_u<2, true>(
        x, r0, r1, log2_ty(x.Y / 2),
        [&](int64_t off, int z, int xx, int yy) {
          Vec4 v;
          v.a = ldg16(xb + off);
          v.b = ldg16(xb + off + 8);
          const int64_t so = s2d_offset(dy, C8, c8, z, xx, yy);
          v.c = ldg16(db + so);
          v.d = ldg16(db + so + pstride);
          return v;
        },
        [&](int64_t, const Vec4& v) {
          accum(v.a, v.c);
          accum(v.b, v.d);
        });
  } else {
    rows_foreach_u<4>(
        x, r0, r1, log2_ty(x.Y),
        [&](int64_t off, int, int, int) {
          Vec2 v;
          v.a = ldg16(xb + off);
          v.b = ldg16(db + off);
          return v;
        },
        [&](int64_t, const Vec2& v) { accum(v.a, v.b); });
  }
  block_reduce<16>(acc, sh, partial + (((size_t)n * x.C8 + c8) * kSlabs + slab) * 16);
}

// dgamma[c] (+)= sum_n red[n][c][1]; dbeta[c] (+)= sum_n red[n][c][0]
__global__ void gn_param_grad_kernel(const float* __restrict__ red, int N, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0, b = 0;
  for (int n = 0; n < N; ++n) {
    a += red[((size_t)n * C + c) * 2];
    b += red[((size_t)n * C + c) * 2 + 1];
  }
  dbeta[c] = accumulate ? dbeta[c] + (float)a : (float)a;
  dgamma[c] = accumulate ? dgamma[c] + (float)b : (float)b;
}

// dx (=|+=) [x>0] * (rstd * (gamma*dy - (s1 + xhat*s2)/m) + add)
// `add` (optional, x's geometry): a second gradient flowing into the same tensor — the residual / fuse-sum pass-through —
// folded into this pass instead of costing a grad_add pass (3 tensor passes) of its own.
template <bool S2D, bool ADD>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(P8 x, P8 dy, int C, int G, const float* __restrict__ stats,
                                                           const float* __restrict__ red, const float* __restrict__ gamma,
                                                           P8 dx, int accumulate, int relu_mask, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int accumulate_params, P8 add) {
  __shared__ float s_k[5 * 8];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int cpg = C / G;
  // the parameter gradients of this chunk (sum over samples of `red`, same order as gn_param_grad_kernel) ride along in
  // one block instead of costing a launch of their own
  if (dgamma && blockIdx.x == 0 && n == 0 && threadIdx.x >= 32 && threadIdx.x < 40) {
    const int c = c8 * 8 + (int)threadIdx.x - 32;
    if (c < C) {
      double a = 0, b = 0;
      for (int nn = 0; nn < x.N; ++nn) {
        a += red[((size_t)nn * C + c) * 2];
        b += red[((size_t)nn * C + c) * 2 + 1];
      }
      dbeta[c] = accumulate_params ? dbeta[c] + (float)a : (float)a;
      dgamma[c] = accumulate_params ? dgamma[c] + (float)b : (float)b;
    }
  }
  if (threadIdx.x < 8) {  // per-(sample, chunk) constants, computed once per block
    const int64_t V = (int64_t)x.Z * x.X * x.Y;
    const float inv_m = 1.0f / ((float)V * (float)cpg);
    const int c = min(c8 * 8 + (int)threadIdx.x, C - 1);
    const int g = c / cpg;
    float s1 = 0.f, s2 = 0.f;
    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
      s1 += gamma[cc] * red[((size_t)n * C + cc) * 2];
      s2 += gamma[cc] * red[((size_t)n * C + cc) * 2 + 1];
    }
    s_k[threadIdx.x] = stats[((size_t)n * G + g) * 2];
    s_k[8 + threadIdx.x] = stats[((size_t)n * G + g) * 2 + 1];
    s_k[16 + threadIdx.x] = (c8 * 8 + (int)threadIdx.x < C) ? gamma[c] : 0.f;
    s_k[24 + threadIdx.x] = s1 * inv_m;
    s_k[32 + threadIdx.x] = s2 * inv_m;
  }
  __syncthreads();
  float mean[8], rstd[8], ga[8], k1[8], k2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = s_k[i]; rstd[i] = s_k[8 + i]; ga[i] = s_k[16 + i]; k1[i] = s_k[24 + i]; k2[i] = s_k[32 + i];
  }
  const int R = x.Z * x.X;
  const int r0 = (int)((int64_t)R * blockIdx.x / gridDim.x), r1 = (int)((int64_t)R * (blockIdx.x + 1) / gridDim.x);
  const bf16* xb = x.ptr + n * x.n_stride + c8 * x.c_stride;
  const bf16* db = dy.ptr + n * dy.n_stride + (S2D ? 0 : c8 * dy.c_stride);
  bf16* ob = dx.ptr + n * dx.n_stride + c8 * dx.c_stride;
  const bf16* ab = ADD ? add.ptr + n * add.n_stride + c8 * add.c_stride : nullptr;
  const int C8 = (int)gridDim.y;
  auto emit = [&](int64_t off, const uint4& xv, const uint4& dv, const uint4& old, const uint4& av) {
    float f[8], d[8], o[8], a[8];
    unpack8(xv, f);
    unpack8(dv, d);
    if constexpr (ADD) unpack8(av, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (f[i] - mean[i]) * rstd[i];
      o[i] = rstd[i] * (ga[i] * d[i] - k1[i] - xh * k2[i]);
      if constexpr (ADD) o[i] += a[i];
      if (relu_mask && !(f[i] > 0.f)) o[i] = 0.f;
    }
    if (accumulate) {
      float p[8];
      unpack8(old, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += p[i];
    }
    stg16(ob + off, pack8(o));
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  if constexpr (S2D) {
    const int64_t pstride = (int64_t)C8 * dy.c_stride;
    rows_foreach_u<2, true>(
        x, r0, r1, log2_ty(x.Y / 2),
        [&](int64_t off, int z, int xx, int yy) {
          Vec8 v;
          v.a = ldg16(xb + off);
          v.b = ldg16(xb + off + 8);
          const int64_t so = s2d_offset(dy, C8, c8, z, xx, yy);
          v.c = ldg16(db + so);
          v.d = ldg16(db + so + pstride);
          v.e = accumulate ? *reinterpret_cast<const uint4*>(ob + off) : zero4;
          v.f = accumulate ? *reinterpret_cast<const uint4*>(ob + off + 8) : zero4;
          if constexpr (ADD) {
            v.g = ldg16(ab + off);
            v.h = ldg16(ab + off + 8);
          }
          return v;
        },
        [&](int64_t off, const Vec8& v) {
          emit(off, v.a, v.c, v.e, v.g);
          emit(off + 8, v.b, v.d, v.f, v.h);
        });
  } else {
    rows_foreach_u<4>(
        x, r0, r1, log2_ty(x.Y),
        [&](int64_t off, int, int, int) {
          Vec4 v;
          v.a = ldg16(xb + off);
          v.b = ldg16(db + off);
          v.c = accumulate ? *reinterpret_cast<const uint4*>(ob + off) : zero4;
          if constexpr (ADD) v.d = ldg16(ab + off);
          return v;
        },
        [&](int64_t off, const Vec4& v) { emit(off, v.a, v.b, v.c, v.d); });
  }
}

// blocks along the row dimension for the elementwise kernels: every thread gets >= ~8 vectors
int ew_blocks(int64_t V) {
  int64_t b = (V + 2047) / 2048;
  return (int)(b > 592 ? 592 : (b < 1 ? 1 : b));
}
}  // namespace

extern "C" int64_t rtp_gn_workspace_bytes(int32_t N, int32_t C8) { return (int64_t)N * C8 * kSlabs * 16 * 4; }

extern "C" int rtp_gn_sums(rtp_p8 x, int32_t C, float* sums, float* workspace, void* stream) {
  RTP_CHECK_ARG(x.ptr && sums && workspace && C > 0 && C <= x.C8 * 8, "rtp_gn_sums: bad args");
  P8 t(x);
  t.C8 = ceil_div(C, 8);
  gn_sums_partial_kernel<<<dim3(kSlabs, t.C8, t.N), 256, 0, (cudaStream_t)stream>>>(t, workspace);
  gn_sums_final_kernel<<<ceil_div(t.N * t.C8 * 8, 128), 128, 0, (cudaStream_t)stream>>>(workspace, t.N, t.C8, C, sums);
  RTP_LAUNCH_CHECK();
}

namespace {
// slab partials -> (mean, rstd) per (sample, group) in one step: thread c sums its channel's slabs (fp64, fixed order),
// then thread g combines the channels of its group — gn_sums_final + gn_finalize without the second launch
__global__ void __launch_bounds__(256) gn_stats_final_kernel(const float* __restrict__ partial, int C8, int C, int G, double count,
                                                             float eps, float* __restrict__ stats) {
  __shared__ double ss[256], sq[256];
  const int n = blockIdx.x, c = threadIdx.x;
  if (c < C) {
    const float* p = partial + ((size_t)n * C8 + (c >> 3)) * kSlabs * 16;
    const int j = c & 7;
    double s = 0, q = 0;
    for (int k = 0; k < kSlabs; ++k) {
      s += p[k * 16 + j];
      q += p[k * 16 + 8 + j];
    }
    ss[c] = (double)(float)s;  // rounded like the fp32 `sums` the two-kernel path hands over
    sq[c] = (double)(float)q;
  }
  __syncthreads();
  if (c < G) {
    const int cpg = C / G;
    double s = 0, q = 0;
    for (int cc = c * cpg; cc < (c + 1) * cpg; ++cc) {
      s += ss[cc];
      q += sq[cc];
    }
    const double m = count * cpg, mean = s / m;
    double var = q / m - mean * mean;
    if (var < 0) var = 0;
    stats[((size_t)n * G + c) * 2] = (float)mean;
    stats[((size_t)n * G + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}
}  // namespace

extern "C" int rtp_gn_stats(rtp_p8 x, int32_t C, int32_t G, float eps, float* stats, float* workspace, void* stream) {
  RTP_CHECK_ARG(x.ptr && stats && workspace && C > 0 && C <= x.C8 * 8 && C <= 256 && G > 0 && C % G == 0, "rtp_gn_stats: bad args");
  P8 t(x);
  t.C8 = ceil_div(C, 8);
  gn_sums_partial_kernel<<<dim3(kSlabs, t.C8, t.N), 256, 0, (cudaStream_t)stream>>>(t, workspace);
  gn_stats_final_kernel<<<t.N, 256, 0, (cudaStream_t)stream>>>(workspace, t.C8, C, G, (double)((int64_t)x.Z * x.X * x.Y), eps, stats);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_gn_finalize(const float* sums, int32_t N, int32_t C, int32_t G, int64_t voxels, float eps, float* stats,
                               void* stream) {
  RTP_CHECK_ARG(sums && stats && N > 0 && C > 0 && G > 0 && C % G == 0 && voxels > 0, "rtp_gn_finalize: bad args");
  gn_finalize_kernel<<<ceil_div(N * G, 64), 64, 0, (cudaStream_t)stream>>>(sums, N, C, G, (double)voxels, eps, stats);
  RTP_LAUNCH_CHECK();
}

namespace {
// `v` must be the space-to-depth view of a tensor with x's grid and ceil(C/8) chunks
bool s2d_view_of(const rtp_p8& v, const rtp_p8& x, int C) {
  return x.Z % 2 == 0 && x.X % 2 == 0 && x.Y % 2 == 0 && v.N == x.N && v.Z == x.Z / 2 && v.X == x.X / 2 && v.Y == x.Y / 2 &&
         v.C8 >= 8 * ceil_div(C, 8);
}

int gn_apply_impl(rtp_p8 x, int32_t C, int32_t G, const float* stats, const float* gamma, const float* beta, rtp_p8 y, bool s2d,
                  void* stream) {
  RTP_CHECK_ARG(x.ptr && y.ptr && stats && gamma && beta, "rtp_gn_apply: null argument");
  RTP_CHECK_ARG(C > 0 && C % G == 0 && C <= x.C8 * 8, "rtp_gn_apply: bad C/G");
  if (s2d)
    RTP_CHECK_ARG(s2d_view_of(y, x, C), "rtp_gn_apply_s2d: y must be the half-resolution view with 8x the chunks (even extents)");
  else
    RTP_CHECK_ARG(C <= y.C8 * 8 && x.N == y.N && x.Z == y.Z && x.X == y.X && x.Y == y.Y, "rtp_gn_apply: geometry mismatch");
  P8 tx(x), ty(y);
  const int64_t V = (int64_t)x.Z * x.X * x.Y;
  const dim3 grid(ew_blocks(V), ceil_div(C, 8), x.N);
  if (s2d)
    gn_apply_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(tx, C, G, stats, gamma, beta, ty);
  else
    gn_apply_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(tx, C, G, stats, gamma, beta, ty);
  RTP_LAUNCH_CHECK();
}

int gn_bwd_reduce_impl(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, float* red, float* workspace, bool s2d,
                       void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && stats && red && workspace, "rtp_gn_bwd_reduce: null argument");
  RTP_CHECK_ARG(C > 0 && C % G == 0 && C <= x.C8 * 8, "rtp_gn_bwd_reduce: bad C/G");
  if (s2d)
    RTP_CHECK_ARG(s2d_view_of(dy, x, C), "rtp_gn_bwd_reduce_s2d: dy must be the space-to-depth view of x's grid");
  else
    RTP_CHECK_ARG(C <= dy.C8 * 8 && x.N == dy.N && x.Z == dy.Z && x.X == dy.X && x.Y == dy.Y, "rtp_gn_bwd_reduce: geometry mismatch");
  P8 tx(x), td(dy);
  tx.C8 = ceil_div(C, 8);
  const dim3 grid(kSlabs, tx.C8, tx.N);
  if (s2d)
    gn_bwd_partial_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(tx, td, C, G, stats, workspace);
  else
    gn_bwd_partial_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(tx, td, C, G, stats, workspace);
  gn_sums_final_kernel<<<ceil_div(tx.N * tx.C8 * 8, 128), 128, 0, (cudaStream_t)stream>>>(workspace, tx.N, tx.C8, C, red);
  RTP_LAUNCH_CHECK();
}

int gn_bwd_apply_impl(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, const float* red, const float* gamma,
                      float* dgamma, float* dbeta, int32_t accumulate_params, rtp_p8 dx, int32_t accumulate_dx, int32_t relu_mask,
                      rtp_p8 add, bool s2d, void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && stats && red && gamma, "rtp_gn_bwd_apply: null argument");
  RTP_CHECK_ARG(C > 0 && C % G == 0 && C <= x.C8 * 8, "rtp_gn_bwd_apply: bad C/G");
  if (s2d)
    RTP_CHECK_ARG(s2d_view_of(dy, x, C), "rtp_gn_bwd_apply_s2d: dy must be the space-to-depth view of x's grid");
  else
    RTP_CHECK_ARG(C <= dy.C8 * 8, "rtp_gn_bwd_apply: bad C/G");
  if (dgamma && dbeta && !dx.ptr)
    gn_param_grad_kernel<<<ceil_div(C, 64), 64, 0, (cudaStream_t)stream>>>(red, x.N, C, dgamma, dbeta, accumulate_params);
  if (dx.ptr) {
    float* dg = (dgamma && dbeta) ? dgamma : nullptr;
    RTP_CHECK_ARG(x.N == dx.N && x.Z == dx.Z && x.X == dx.X && x.Y == dx.Y && C <= dx.C8 * 8, "rtp_gn_bwd_apply: dx geometry mismatch");
    if (add.ptr)
      RTP_CHECK_ARG(x.N == add.N && x.Z == add.Z && x.X == add.X && x.Y == add.Y && C <= add.C8 * 8, "rtp_gn_bwd_apply: add geometry mismatch");
    P8 tx(x), td(dy), to(dx), ta(add);
    const int64_t V = (int64_t)x.Z * x.X * x.Y;
    const dim3 grid(ew_blocks(V), ceil_div(C, 8), x.N);
    auto kern = s2d ? (add.ptr ? gn_bwd_apply_kernel<true, true> : gn_bwd_apply_kernel<true, false>)
                    : (add.ptr ? gn_bwd_apply_kernel<false, true> : gn_bwd_apply_kernel<false, false>);
    kern<<<grid, 256, 0, (cudaStream_t)stream>>>(tx, td, C, G, stats, red, gamma, to, accumulate_dx, relu_mask, dg, dbeta,
                                                 accumulate_params, ta);
  } else {
    RTP_CHECK_ARG(!add.ptr, "rtp_gn_bwd_apply: add given without dx");
  }
  RTP_LAUNCH_CHECK();
}
}  // namespace

extern "C" int rtp_gn_apply(rtp_p8 x, int32_t C, int32_t G, const float* stats, const float* gamma, const float* beta,
                            rtp_p8 y, void* stream) {
  return gn_apply_impl(x, C, G, stats, gamma, beta, y, false, stream);
}
extern "C" int rtp_gn_apply_s2d(rtp_p8 x, int32_t C, int32_t G, const float* stats, const float* gamma, const float* beta,
                                rtp_p8 y_s2d, void* stream) {
  return gn_apply_impl(x, C, G, stats, gamma, beta, y_s2d, true, stream);
}
extern "C" int rtp_gn_bwd_reduce(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, float* red,
                                 float* workspace, void* stream) {
  return gn_bwd_reduce_impl(x, dy, C, G, stats, red, workspace, false, stream);
}
extern "C" int rtp_gn_bwd_reduce_s2d(rtp_p8 x, rtp_p8 dy_s2d, int32_t C, int32_t G, const float* stats, float* red,
                                     float* workspace, void* stream) {
  return gn_bwd_reduce_impl(x, dy_s2d, C, G, stats, red, workspace, true, stream);
}
extern "C" int rtp_gn_bwd_apply(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, const float* red,
                                const float* gamma, float* dgamma, float* dbeta, int32_t accumulate_params, rtp_p8 dx,
                                int32_t accumulate_dx, int32_t relu_mask, rtp_p8 add, void* stream) {
  return gn_bwd_apply_impl(x, dy, C, G, stats, red, gamma, dgamma, dbeta, accumulate_params, dx, accumulate_dx, relu_mask, add,
                           false, stream);
}
extern "C" int rtp_gn_bwd_apply_s2d(rtp_p8 x, rtp_p8 dy_s2d, int32_t C, int32_t G, const float* stats, const float* red,
                                    const float* gamma, float* dgamma, float* dbeta, int32_t accumulate_params, rtp_p8 dx,
                                    int32_t accumulate_dx, int32_t relu_mask, rtp_p8 add, void* stream) {
  return gn_bwd_apply_impl(x, dy_s2d, C, G, stats, red, gamma, dgamma, dbeta, accumulate_params, dx, accumulate_dx, relu_mask, add,
                           true, stream);
}

// ================================================================================================ stem / bias grads
namespace {

// y[c] = w[c] * x + b[c], x = channel 0 of a P8 tensor (single-channel radar cube)
__global__ void __launch_bounds__(256) stem_fwd_kernel(P8 x, const float* __restrict__ w, const float* __restrict__ b, int C,
                                                       P8 y) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  float wv[8], bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c8 * 8 + i;
    wv[i] = c < C ? w[c] : 0.f;
    bv[i] = c < C ? b[c] : 0.f;
  }
  const int64_t V = (int64_t)x.Z * x.X * x.Y;
  const bf16* xb = x.ptr + n * x.n_stride;
  bf16* yb = y.ptr + n * y.n_stride + c8 * y.c_stride;
  for (int64_t v = blockIdx.x * 256ll + threadIdx.x; v < V; v += (int64_t)gridDim.x * 256) {
    const int64_t off = voxel_off(x, v);
    const float xv = __bfloat162float(xb[off]);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(wv[i], xv, bv[i]);
    stg16(yb + off, pack8(f));
  }
}

// partial[n][c8][slab][16] = (sum dy[0..7], sum dy[0..7] * x)
__global__ void __launch_bounds__(256) stem_bwd_partial_kernel(P8 x, P8 dy, float* __restrict__ partial) {
  __shared__ float sh[8 * 16];
  const int slab = blockIdx.x, c8 = blockIdx.y, n = blockIdx.z;
  const int64_t V = (int64_t)x.Z * x.X * x.Y;
  const int64_t v0 = V * slab / kSlabs, v1 = V * (slab + 1) / kSlabs;
  const bf16* xb = x.ptr + n * x.n_stride;
  const bf16* db = dy.ptr + n * dy.n_stride + c8 * dy.c_stride;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {
    const int64_t off = voxel_off(x, v);
    const float xv = __bfloat162float(xb[off]);
    float d[8];
    unpack8(ldg16(db + off), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += d[i];
      acc[8 + i] += d[i] * xv;
    }
  }
  block_reduce<16>(acc, sh, partial + (((size_t)n * gridDim.y + c8) * kSlabs + slab) * 16);
}

// out0[c] (+)= sum_{n,slab} partial[..][j], out1[c] (+)= sum partial[..][8+j]: one block per 8-channel chunk, the N*kSlabs
// partial rows spread over the threads, fixed-shape tree reduction (deterministic).  (One thread per channel walking
// 2*N*kSlabs dependent loads took 47 us.)
__global__ void __launch_bounds__(256) slab_final_kernel(const float* __restrict__ partial, int N, int C8, int C,
                                                         float* __restrict__ out0, float* __restrict__ out1, int accumulate) {
  __shared__ float sh[8 * 16];
  __shared__ float res[16];
  const int c8 = blockIdx.x;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int i = threadIdx.x; i < N * kSlabs; i += 256) {
    const int n = i / kSlabs, k = i - n * kSlabs;
    const float4* p = reinterpret_cast<const float4*>(partial + (((size_t)n * C8 + c8) * kSlabs + k) * 16);
    const float4 a = p[0], b = p[1], c = p[2], d = p[3];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    acc[8] += c.x; acc[9] += c.y; acc[10] += c.z; acc[11] += c.w; acc[12] += d.x; acc[13] += d.y; acc[14] += d.z; acc[15] += d.w;
  }
  block_reduce<16>(acc, sh, res);
  __syncthreads();
  if (threadIdx.x < 8) {
    const int c = c8 * 8 + threadIdx.x;
    if (c < C) {
      if (out0) out0[c] = accumulate ? out0[c] + res[threadIdx.x] : res[threadIdx.x];
      if (out1) out1[c] = accumulate ? out1[c] + res[8 + threadIdx.x] : res[8 + threadIdx.x];
    }
  }
}

}  // namespace

extern "C" int rtp_stem_fwd(rtp_p8 x, const float* w, const float* b, int32_t C, rtp_p8 y, void* stream) {
  RTP_CHECK_ARG(x.ptr && y.ptr && w && b && C > 0 && C <= y.C8 * 8, "rtp_stem_fwd: bad args");
  RTP_CHECK_ARG(x.N == y.N && x.Z == y.Z && x.X == y.X && x.Y == y.Y, "rtp_stem_fwd: geometry mismatch");
  const int64_t V = (int64_t)x.Z * x.X * x.Y;
  stem_fwd_kernel<<<dim3(ew_blocks(V), ceil_div(C, 8), x.N), 256, 0, (cudaStream_t)stream>>>(P8(x), w, b, C, P8(y));
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_stem_bwd(rtp_p8 x, rtp_p8 dy, int32_t C, float* dw, float* db, int32_t accumulate, float* workspace,
                            void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && dw && db && workspace && C > 0 && C <= dy.C8 * 8, "rtp_stem_bwd: bad args");
  const int C8 = ceil_div(C, 8);
  stem_bwd_partial_kernel<<<dim3(kSlabs, C8, x.N), 256, 0, (cudaStream_t)stream>>>(P8(x), P8(dy), workspace);
  slab_final_kernel<<<C8, 256, 0, (cudaStream_t)stream>>>(workspace, x.N, C8, C, db, dw, accumulate);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_channel_sum(rtp_p8 x, int32_t C, float* out, int32_t accumulate, float* workspace, void* stream) {
  RTP_CHECK_ARG(x.ptr && out && workspace && C > 0 && C <= x.C8 * 8, "rtp_channel_sum: bad args");
  P8 t(x);
  t.C8 = ceil_div(C, 8);
  gn_sums_partial_kernel<<<dim3(kSlabs, t.C8, t.N), 256, 0, (cudaStream_t)stream>>>(t, workspace);
  slab_final_kernel<<<t.C8, 256, 0, (cudaStream_t)stream>>>(workspace, t.N, t.C8, C, out, nullptr, accumulate);
  RTP_LAUNCH_CHECK();
}

// Targeted L1 / shared-memory preference (see rtp_set_shared_carveout, layout.cu): the streaming GroupNorm kernels get the
// maximum-shared split so that they can become resident beside the persistent weight-gradient / conv CTAs.
int rtp_norm_set_carveout(int pct, int backward_only) {
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* f) {
    const cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (r != cudaSuccess) e = r;
  };
  if (!backward_only) {
    set((const void*)gn_sums_partial_kernel);
    set((const void*)gn_apply_kernel<false>);
    set((const void*)gn_apply_kernel<true>);
    set((const void*)gn_bwd_partial_kernel<false>);
    set((const void*)gn_bwd_partial_kernel<true>);
  }
  set((const void*)gn_sums_final_kernel);
  set((const void*)gn_finalize_kernel);
  set((const void*)gn_param_grad_kernel);
  set((const void*)gn_bwd_apply_kernel<false, false>);
  set((const void*)gn_bwd_apply_kernel<false, true>);
  set((const void*)gn_bwd_apply_kernel<true, false>);
  set((const void*)gn_bwd_apply_kernel<true, true>);
  set((const void*)gn_stats_final_kernel);
  set((const void*)slab_final_kernel);
  return (int)e;
}
