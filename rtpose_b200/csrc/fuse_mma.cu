// fuse_mma.cu — the branch-exchange sum with the y interpolation of the low-resolution terms on the tensor cores:
//
//   out = [relu]( sum_i same[i] + sum_j trilinear_up(low[j]) + bias )        (hr_util/hr3d.py:213-227, align_corners=True)
//
// fuse_sum_tile_kernel (fuse.cu) spends 13 warp instructions per 16-byte output vector on fp32 blends and index arithmetic
// (ncu: 72 % issue utilisation, 0.34 of the HBM roofline).  Trilinear interpolation is separable; here a WARP walks the
// output rows x of an x segment of one (sample, chunk, z) plane:
//   * z: the two low-resolution planes of z are blended once per low-resolution row xl and kept in registers (lane = yl) for
//     the output rows that use it — an "even" and an "odd" register set by the parity of xl, the raw vectors of the next xl
//     prefetched one advance ahead;
//   * x: per output row, w0 * Z[xl] + w1 * Z[xl + 1] (8 channels per lane), rounded to bf16 and written as one 16-byte row
//     [yl][8 channels] of the warp's shared-memory scratch — exactly the B operand layout ldmatrix.trans wants;
//   * y: [Y x Yl] x [Yl x 8 channels] on mma.sync.m16n8k16, the interpolation matrix (bf16 high + low parts, 16 mantissa
//     bits) as A fragments in shared memory; all terms accumulate into the same fp32 fragments, which start from the sum of
//     the same-resolution terms (4-byte loads in the accumulator layout: a warp reads whole 128-byte lines, issued one row
//     ahead) and leave as 4-byte stores of whole lines.
// ~4 warp instructions per output vector — but see rtp_fuse_sum_mma below: opt-in, measured slower.  The blended low-resolution row is rounded to bf16 before the y interpolation
// (one rounding more than the CUDA-core kernel; the reference under autocast rounds every interpolated tensor and every
// partial sum to bf16).  Fixed summation order (deterministic).
//
// Roofline: HBM — same[i] read once, out written once; the low-resolution terms are L2-resident.
#include <cstdlib>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kWarps = 4;
constexpr int kSegRows = 16;  // output rows per unit

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
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// D[y][c] += Wy[y][yl] * B[yl][c] for one low term: B fragments by ldmatrix.trans from the warp's scratch rows, A fragments
// (high and low weight parts) from the CTA's table
template <int MT, int KSJ>
__device__ __forceinline__ void term_mma(float (&d)[MT][4], const uint4* af, uint32_t rows_s, int lane) {
  uint32_t b[4];
  ldmatrix_x4_trans(rows_s + (uint32_t)(lane & (KSJ * 16 - 1)) * 16u, b);  // KSJ = 1: lanes 16-31 re-read rows 0-15
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KSJ; ++ks) {
      const uint4 ahi = af[((mt * KSJ + ks) * 2) * 32], alo = af[((mt * KSJ + ks) * 2 + 1) * 32];
      mma_bf16(d[mt], ahi, b[2 * ks], b[2 * ks + 1]);
      mma_bf16(d[mt], alo, b[2 * ks], b[2 * ks + 1]);
    }
}

struct FMK {
  P8 out, same[4], low[3];
  const float* bias;
  int C, C8, n_same, n_low, relu;
  int nseg, nunits;
  uint32_t a_off[3];  // uint4 index of term j's A fragments: [mt][ks][hi | lo][32 lanes]
  uint32_t b_off[3];  // byte offset of term j's B rows inside a warp's scratch: [16 * KS_j rows][16 bytes]
  uint32_t a_total;   // uint4s of all A fragments
  uint32_t b_bytes;   // bytes of one warp's scratch
};

// MT = Y / 16 row blocks of the output row; KS0, KS1, KS2 = k16 steps of the three low terms (0: term absent)
template <int MT, int KS0, int KS1, int KS2>
__global__ void __launch_bounds__(kWarps * 32, 3) fuse_sum_mma_kernel(const __grid_constant__ FMK p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NLOW = (KS0 > 0) + (KS1 > 0) + (KS2 > 0);
  auto ksj = [](int j) { return j == 0 ? KS0 : (j == 1 ? KS1 : KS2); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const P8& o = p.out;
  const int Y = o.Y, X = o.X;
  uint4* afr = reinterpret_cast<uint4*>(smem);
  uint8_t* brow = smem + (size_t)p.a_total * 16 + (size_t)warp * p.b_bytes;
  const uint32_t brow_s = smem_u32(brow);

  // ---- A fragments of the interpolation matrices Wy_j[y][yl] (forward weights), bf16 high and low parts; scratch rows zeroed
#pragma unroll
  for (int j = 0; j < NLOW; ++j) {
    const int Yl = p.low[j].Y;
    const float sy = ac_scale(Yl, Y);
    const int KSJ = ksj(j);
    for (int e = warp; e < MT * KSJ; e += kWarps) {
      const int mt = e / KSJ, ks = e - mt * KSJ;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int y = mt * 16 + g + (i & 1) * 8;
        const int k0 = ks * 16 + 2 * q + (i >> 1) * 8;
        const Axis ay = ac_axis(y, Yl, sy);
        float w[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int yl = k0 + t;
          w[t] = (yl == ay.i0 ? ay.w0 : 0.f) + (yl == ay.i1 ? ay.w1 : 0.f);
        }
        hi[i] = pack_bf16x2(w[0], w[1]);
        const float2 hf = unpack_bf16x2(hi[i]);
        lo[i] = pack_bf16x2(w[0] - hf.x, w[1] - hf.y);
      }
      afr[p.a_off[j] + (uint32_t)((mt * KSJ + ks) * 2) * 32 + lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      afr[p.a_off[j] + (uint32_t)((mt * KSJ + ks) * 2 + 1) * 32 + lane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  for (uint32_t i = lane * 16; i < p.b_bytes; i += 32 * 16) *reinterpret_cast<uint4*>(brow + i) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  float bias[2] = {0.f, 0.f};
  const int wglobal = blockIdx.x * kWarps + warp, wtotal = gridDim.x * kWarps;
  const int64_t xstride = (int64_t)o.Yp * 8;

  for (int t = wglobal; t < p.nunits; t += wtotal) {
    const int seg = t % p.nseg;
    int r = t / p.nseg;
    const int z = r % o.Z;
    r /= o.Z;
    const int c8 = r % p.C8, n = r / p.C8;
    const int xa = seg * kSegRows, xb = min(X, xa + kSegRows);
    if (p.bias) {
      bias[0] = c8 * 8 + 2 * q < p.C ? p.bias[c8 * 8 + 2 * q] : 0.f;
      bias[1] = c8 * 8 + 2 * q + 1 < p.C ? p.bias[c8 * 8 + 2 * q + 1] : 0.f;
    }
    // ---- per-term state: lane = yl; Z-blended rows of xl = a (set by parity of a) and a + 1, raw vectors of a + 2 in flight
    float zE[NLOW][8], zO[NLOW][8];
    uint4 pre0[NLOW], pre1[NLOW];
    int a[NLOW];
    const bf16* lrow0[NLOW];  // (z0, x = 0, yl) and (z1, x = 0, yl) of this lane
    const bf16* lrow1[NLOW];
    float wz0[NLOW], wz1[NLOW], sx[NLOW];
    int64_t lxs[NLOW];
    auto zblend = [&](int j, const uint4& r0, const uint4& r1, float (&dst)[8]) {
      float f0[8], f1[8];
      unpack8(r0, f0);
      unpack8(r1, f1);
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[e] = wz0[j] * f0[e] + wz1[j] * f1[e];
    };
    auto fetch = [&](int j, int xl, uint4& r0, uint4& r1) {
      const P8& l = p.low[j];
      if (lane < l.Y && xl < l.X) {
        r0 = ldg16(lrow0[j] + (int64_t)xl * lxs[j]);
        r1 = ldg16(lrow1[j] + (int64_t)xl * lxs[j]);
      } else {
        r0 = r1 = make_uint4(0u, 0u, 0u, 0u);
      }
    };
#pragma unroll
    for (int j = 0; j < NLOW; ++j) {
      const P8& l = p.low[j];
      const Axis az = ac_axis(z, l.Z, ac_scale(l.Z, o.Z));
      const bf16* lb = l.ptr + (int64_t)n * l.n_stride + (int64_t)c8 * l.c_stride;
      const int yl = lane < l.Y ? lane : 0;
      lrow0[j] = lb + l.voxel(az.i0, 0, yl);
      lrow1[j] = lb + l.voxel(az.i1, 0, yl);
      lxs[j] = (int64_t)l.Yp * 8;
      wz0[j] = az.w0; wz1[j] = az.w1;
      sx[j] = ac_scale(l.X, X);
      a[j] = (int)(sx[j] * (float)xa);
      uint4 r0, r1;
      fetch(j, a[j], r0, r1);
      fetch(j, a[j] + 1, pre0[j], pre1[j]);
      if (a[j] & 1) { zblend(j, r0, r1, zO[j]); zblend(j, pre0[j], pre1[j], zE[j]); }
      else { zblend(j, r0, r1, zE[j]); zblend(j, pre0[j], pre1[j], zO[j]); }
      fetch(j, a[j] + 2, pre0[j], pre1[j]);
    }
    // ---- same-resolution terms of the first row, in the accumulator layout: (y = mt*16 + g [+ 8], channels 2q, 2q + 1)
    const int64_t obase = (int64_t)n * o.n_stride + (int64_t)c8 * o.c_stride + o.voxel(z, xa, 0) + 2 * q;
    uint32_t sraw[4][MT][2];
    auto load_same = [&](int x) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s >= p.n_same) break;
        const bf16* sb = p.same[s].ptr + (int64_t)n * p.same[s].n_stride + (int64_t)c8 * p.same[s].c_stride + o.voxel(z, x, 0) + 2 * q;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) sraw[s][mt][h] = __ldg(reinterpret_cast<const uint32_t*>(sb + (mt * 16 + g + h * 8) * 8));
      }
    };
    load_same(xa);

    for (int x = xa; x < xb; ++x) {
      // ---- x interpolation of every term into the scratch rows (bf16), advancing the z-blended sets as xl moves
#pragma unroll
      for (int j = 0; j < NLOW; ++j) {
        const P8& l = p.low[j];
        const Axis ax = ac_axis(x, l.X, sx[j]);
        if (ax.i0 != a[j]) {  // i0 == a + 1: xl = a + 2 replaces xl = a in the set of a's parity
          if (a[j] & 1) zblend(j, pre0[j], pre1[j], zO[j]); else zblend(j, pre0[j], pre1[j], zE[j]);
          a[j] = ax.i0;
          fetch(j, a[j] + 2, pre0[j], pre1[j]);
        }
        const float wa = ax.i1 == ax.i0 ? ax.w0 + ax.w1 : ax.w0, wb = ax.i1 == ax.i0 ? 0.f : ax.w1;
        const float wE = (a[j] & 1) ? wb : wa, wO = (a[j] & 1) ? wa : wb;
        if (lane < l.Y) {
          uint4 v;
          v.x = pack_bf16x2(wE * zE[j][0] + wO * zO[j][0], wE * zE[j][1] + wO * zO[j][1]);
          v.y = pack_bf16x2(wE * zE[j][2] + wO * zO[j][2], wE * zE[j][3] + wO * zO[j][3]);
          v.z = pack_bf16x2(wE * zE[j][4] + wO * zO[j][4], wE * zE[j][5] + wO * zO[j][5]);
          v.w = pack_bf16x2(wE * zE[j][6] + wO * zO[j][6], wE * zE[j][7] + wO * zO[j][7]);
          *reinterpret_cast<uint4*>(brow + p.b_off[j] + lane * 16) = v;
        }
      }
      // ---- accumulators start from the same-resolution sum (loaded one row ahead); then request the next row
      float d[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[mt][i] = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s >= p.n_same) break;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 f = unpack_bf16x2(sraw[s][mt][h]);
            d[mt][2 * h] += f.x;
            d[mt][2 * h + 1] += f.y;
          }
      }
      if (x + 1 < xb) load_same(x + 1);
      __syncwarp();
      // ---- y interpolation: D[y][c] += Wy_j[y][yl] * B_j[yl][c]
      if constexpr (KS0 > 0) term_mma<MT, KS0>(d, afr + p.a_off[0] + lane, brow_s + p.b_off[0], lane);
      if constexpr (KS1 > 0) term_mma<MT, KS1>(d, afr + p.a_off[1] + lane, brow_s + p.b_off[1], lane);
      if constexpr (KS2 > 0) term_mma<MT, KS2>(d, afr + p.a_off[2] + lane, brow_s + p.b_off[2], lane);
      // ---- bias, ReLU, 4-byte stores (a warp writes whole 128-byte lines)
      bf16* orow = o.ptr + obase + (int64_t)(x - xa) * xstride;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v0 = d[mt][2 * h] + bias[0], v1 = d[mt][2 * h + 1] + bias[1];
          if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          *reinterpret_cast<uint32_t*>(orow + (mt * 16 + g + h * 8) * 8) = pack_bf16x2(v0, v1);
        }
      __syncwarp();  // every lane has read the scratch rows before the next row overwrites them
    }
  }
}

template <int MT>
int launch_mt(const FMK& k, int ks0, int ks1, int ks2, int grid, size_t smem, cudaStream_t st) {
  const int code = ks0 * 100 + ks1 * 10 + ks2;
#define RTP_FM_CASE(A, B, C)                                                                                           \
  case A * 100 + B * 10 + C: {                                                                                         \
    static size_t conf[RTP_MAX_DEVICES];                                                                               \
    size_t& c = conf[rtp_current_device()];                                                                            \
    if (smem > c) {                                                                                                    \
      if (cudaFuncSetAttribute(fuse_sum_mma_kernel<MT, A, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
        return -1;                                                                                                     \
      c = smem;                                                                                                        \
    }                                                                                                                  \
    fuse_sum_mma_kernel<MT, A, B, C><<<grid, kWarps * 32, smem, st>>>(k);                                              \
    return 1;                                                                                                          \
  }
  switch (code) {
    RTP_FM_CASE(1, 0, 0)
    RTP_FM_CASE(2, 0, 0)
    RTP_FM_CASE(1, 1, 0)
    RTP_FM_CASE(2, 1, 0)
    RTP_FM_CASE(1, 1, 1)
    RTP_FM_CASE(2, 1, 1)
    default: return 0;
  }
#undef RTP_FM_CASE
}

}  // namespace

// Called by rtp_fuse_sum (fuse.cu).  Returns 1 when the launch was issued, 0 when the shape is not supported (the caller
// falls back to the CUDA-core kernels), < 0 on error.  Terms must be ordered by descending Y (the branch order).
int rtp_fuse_sum_mma(const rtp_fuse_desc* d, void* stream) {
  // Opt-in (RTP_FUSE_MMA=1, read per call so that the tests can switch it): correct, but measured SLOWER than the tile kernel
  // of fuse.cu (32 ch, 3 low terms: 0.209 vs 0.175 ms; 2 low terms: 0.152 vs 0.144 ms).  Unlike the backward direction
  // (upsample_mma.cu), the per-lane state of the z / x blend leaves no registers for the interpolation matrices, and reading
  // their fragments from shared memory costs 32 16-byte loads (128 wavefronts) per output row.
  const char* on = getenv("RTP_FUSE_MMA");
  if (!on || on[0] != '1' || d->n_low < 1) return 0;
  const int Y = d->out.Y;
  if (Y % 16 != 0 || Y > 64 || Y < 16) return 0;
  int ks[3] = {0, 0, 0};
  for (int j = 0; j < d->n_low; ++j) {
    const int Yl = d->low[j].Y;
    if (Yl < 1 || Yl > 32 || Yl > Y) return 0;
    ks[j] = Yl > 16 ? 2 : 1;
    if (j > 0 && ks[j] > ks[j - 1]) return 0;  // instantiated patterns: (2|1, 1|0, 1|0)
  }
  FMK k;
  k.out = P8(d->out);
  k.C = d->C;
  k.C8 = ceil_div(d->C, 8);
  k.n_same = d->n_same; k.n_low = d->n_low; k.relu = d->relu; k.bias = d->bias;
  for (int i = 0; i < 4; ++i) k.same[i] = P8(i < d->n_same ? d->same[i] : d->out);
  for (int j = 0; j < 3; ++j) k.low[j] = P8(j < d->n_low ? d->low[j] : d->out);
  const int MT = Y / 16;
  uint32_t a_total = 0, b_bytes = 0;
  for (int j = 0; j < 3; ++j) {
    k.a_off[j] = a_total;
    k.b_off[j] = b_bytes;
    a_total += (uint32_t)(MT * ks[j] * 2 * 32);
    b_bytes += (uint32_t)(ks[j] * 16 * 16);
  }
  if (b_bytes == 0) return 0;
  k.a_total = a_total;
  k.b_bytes = (b_bytes + 127u) & ~127u;
  k.nseg = ceil_div(d->out.X, kSegRows);
  const int64_t nunits = (int64_t)k.nseg * d->out.Z * k.C8 * d->out.N;
  if (nunits > 0x7fffffff) return 0;
  k.nunits = (int)nunits;
  const size_t smem = (size_t)a_total * 16 + (size_t)kWarps * k.b_bytes;
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int want = ceil_div(k.nunits, kWarps);
  const int grid = want < 3 * nsm ? want : 3 * nsm;
  int r = 0;
  switch (MT) {
    case 1: r = launch_mt<1>(k, ks[0], ks[1], ks[2], grid, smem, (cudaStream_t)stream); break;
    case 2: r = launch_mt<2>(k, ks[0], ks[1], ks[2], grid, smem, (cudaStream_t)stream); break;
    case 3: r = launch_mt<3>(k, ks[0], ks[1], ks[2], grid, smem, (cudaStream_t)stream); break;
    case 4: r = launch_mt<4>(k, ks[0], ks[1], ks[2], grid, smem, (cudaStream_t)stream); break;
    default: r = 0;
  }
  if (r < 0) rtp_set_error("rtp_fuse_sum: cudaFuncSetAttribute failed for the tensor-core kernel");
  return r;
}
