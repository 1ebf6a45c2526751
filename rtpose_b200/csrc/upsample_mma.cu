// upsample_mma.cu — the y and x reductions of the transposed trilinear upsample (align_corners=True) with the y reduction
// on the tensor cores: the backward of every branch exchange (hr_util/hr3d.py:219-229 F.interpolate) and of the final concat.
//
//   t2[n][c][z][xl][yl] = sum_x wx(x -> xl) * sum_y wy(y -> yl) * dout[n][c][z][x][y]
//
// The CUDA-core kernels (fuse.cu: upsample_bwd_yx_kernel / _shfl_kernel) are instruction-bound: ncu counts 5.7 warp
// instructions per 16-byte input vector (index arithmetic, hat weights, bf16 unpacking, 16 FMAs) and 0.31 of the HBM
// roofline.  Here a WARP owns a tile (n, chunk, z, kTXL low-resolution x positions) and streams its full-resolution rows:
//   * the rows of a tile are contiguous in the P8 layout, so ONE 1-D bulk async copy (cp.async.bulk + mbarrier) stages them
//     in the warp's private two-stage shared-memory ring; the copy of the tile after next is in flight while a tile is reduced;
//   * y reduction of a row = a [Yl x Y] x [Y x 8 channels] product: the B fragments come straight out of the staged row with
//     ldmatrix.trans (a P8 row is [y][8 channels], 16 bytes per y), the A fragments — the banded interpolation matrix, split
//     into a bf16 high and low part so that the weights keep 16 mantissa bits — live in registers for the whole kernel,
//     mma.sync.m16n8k16 accumulates in fp32;
//   * the accumulator fragment leaves every thread with (4 yl) x (2 channels) of the y-reduced row, so the x reduction is
//     thread-local: a full-resolution row contributes to at most two low-resolution rows (xl = i0, i0 + 1), which are two
//     rolling accumulator sets; a finished xl row goes out as 4-byte stores that a warp coalesces into whole 128-byte lines.
// ~0.6 warp instructions per input vector.  Sums are formed in a fixed order (deterministic).  Weights: hat_weight() of
// fuse.cu (the forward's w0 / w1 expressions).  No TMEM: the kernel co-resides with the weight-gradient kernels it overlaps.
//
// Roofline: HBM — dout read once (+ halo rows: (kTXL / scale + 2) rows per kTXL outputs), t2 written once.
#include <cstdlib>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ float ac_scale_(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ __forceinline__ float hat_weight_(int d, int l, float scale) {
  const float src = scale * (float)d;
  const float w = src >= (float)l ? 1.f - (src - (float)l) : src - (float)(l - 1);
  return w > 0.f ? w : 0.f;
}
__device__ __forceinline__ void hat_range_(int l, int in_n, float scale, float inv, int& lo, int& hi) {
  lo = max(0, (int)floorf((float)(l - 1) * inv));
  hi = min(in_n - 1, (int)ceilf((float)(l + 1) * inv));
  if (scale * (float)lo <= (float)(l - 1)) ++lo;
  if (scale * (float)hi >= (float)(l + 1)) --hi;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct UBM {
  P8 in, out;
  int C8, ntx, txl, nunits, nks;  // x segments per plane, low-resolution x positions per segment, k16 steps per row
  uint32_t stage_bytes;
};

// chunks of kRows full-resolution rows in flight per warp: 3 for the one-block kernels (96 registers, 2 CTAs per SM); the
// two-block kernel (Yl > 16, 161 registers, issue-bound at 8 warps per SM) runs 3 CTAs per SM with 2 stages (-7 %)
constexpr int stages_of(int nmt) { return nmt == 2 ? 2 : 3; }
constexpr int kRows = 8;

// NMT = 16-row blocks of the low-resolution y extent (1: Yl <= 16, 2: Yl <= 32); NKS = Y / 16 k16 steps per row (Y <= 64)
template <int NMT, int NKS>
__global__ void __launch_bounds__(kWarps * 32) upsample_bwd_yx_mma_kernel(const __grid_constant__ UBM p) {
  constexpr int kStages = stages_of(NMT);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kWarps][kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int Y = p.in.Y, Yl = p.out.Y, X = p.in.X, Xl = p.out.X;
  const float sy = ac_scale_(Yl, Y), sx = ac_scale_(Xl, X);
  const float ix = 1.f / sx;
  uint8_t* ring = smem + (size_t)warp * kStages * p.stage_bytes;
  const uint32_t ring_s = smem_u32(ring);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bar_full[warp][s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  // A fragments of the interpolation matrix Wy[yl][y] (row-major m16 x k16 per (mt, ks)), high and low bf16 parts.  Blocks
  // outside the band of the matrix are all-zero and are multiplied anyway: straight-line code lets the independent
  // accumulation chains of two rows interleave (per-block branches cost more than the four extra MMAs of the 2x case).
  uint32_t ah[NMT][NKS][4], al[NMT][NKS][4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yl = mt * 16 + g + (i & 1) * 8;
        const int d0 = ks * 16 + 2 * q + (i >> 1) * 8;
        float w0 = 0.f, w1 = 0.f;
        if (yl < Yl) {
          w0 = d0 < Y ? hat_weight_(d0, yl, sy) : 0.f;
          w1 = d0 + 1 < Y ? hat_weight_(d0 + 1, yl, sy) : 0.f;
        }
        const uint32_t h = pack_bf16x2(w0, w1);
        const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&h);
        const float2 hf = __bfloat1622float2(hb);
        ah[mt][ks][i] = h;
        al[mt][ks][i] = pack_bf16x2(w0 - hf.x, w1 - hf.y);
      }

  const int wglobal = blockIdx.x * kWarps + warp, wtotal = gridDim.x * kWarps;
  const uint32_t row_bytes = (uint32_t)p.in.Yp * 16u;

  struct Unit {
    int n, c8, z, xl0, xl1, r0, nrows;
  };
  auto unit_geom = [&](int t) {
    Unit u;
    const int xt = t % p.ntx;
    int r = t / p.ntx;
    u.z = r % p.in.Z;
    r /= p.in.Z;
    u.c8 = r % p.C8;
    u.n = r / p.C8;
    u.xl0 = xt * p.txl;
    u.xl1 = min(Xl, u.xl0 + p.txl) - 1;
    int r1, tmp;
    hat_range_(u.xl0, X, sx, ix, u.r0, tmp);
    hat_range_(u.xl1, X, sx, ix, tmp, r1);
    u.nrows = r1 - u.r0 + 1;
    return u;
  };

  // ---- producer cursor (warp-uniform; lane 0 issues): the chunk sequence of this warp's units, kStages chunks ahead
  int pt = wglobal, pc = 0;
  Unit pu = unit_geom(pt < p.nunits ? pt : 0);
  auto produce = [&](int s) {
    if (pt >= p.nunits) return;
    const int row0 = pc * kRows;
    const int rows = min(kRows, pu.nrows - row0);
    if (lane == 0) {
      const bf16* src = p.in.ptr + (int64_t)pu.n * p.in.n_stride + (int64_t)pu.c8 * p.in.c_stride + p.in.voxel(pu.z, pu.r0 + row0, 0);
      const uint32_t bytes = (uint32_t)rows * row_bytes;
      mbar_arrive_expect_tx(&bar_full[warp][s], bytes);
      bulk_g2s(ring + (size_t)s * p.stage_bytes, src, bytes, &bar_full[warp][s]);
    }
    ++pc;
    if (pc * kRows >= pu.nrows) {
      pt += wtotal;
      pc = 0;
      if (pt < p.nunits) pu = unit_geom(pt);
    }
  };
#pragma unroll
  for (int s = 0; s < kStages; ++s) produce(s);

  uint32_t cnt = 0;  // chunks consumed
  for (int t = wglobal; t < p.nunits; t += wtotal) {
    const Unit u = unit_geom(t);
    bf16* out_row0 = p.out.ptr + (int64_t)u.n * p.out.n_stride + (int64_t)u.c8 * p.out.c_stride + p.out.voxel(u.z, 0, 0) + 2 * q;
    const int64_t out_xstride = (int64_t)p.out.Yp * 8;
    // x reduction state: a full-resolution row contributes to the low-resolution rows xl = a and a + 1; they accumulate in
    // the even / odd set according to the parity of xl, so advancing a re-uses the registers without moving them
    float accE[NMT][4], accO[NMT][4];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) accE[mt][i] = accO[mt][i] = 0.f;
    auto flush = [&](int xl, float (&acc)[NMT][4]) {  // store row xl if this unit owns it, then reset the set
      if (xl >= u.xl0 && xl <= u.xl1) {
        bf16* o = out_row0 + (int64_t)xl * out_xstride;
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int yl = mt * 16 + g + h * 8;
            if (yl < Yl) *reinterpret_cast<uint32_t*>(o + yl * 8) = pack_bf16x2(acc[mt][2 * h], acc[mt][2 * h + 1]);
          }
      }
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
    };
    int a = (int)(sx * (float)u.r0);
    // y reduction of one staged row on the tensor cores: separate chains for the high and low weight parts
    auto yreduce = [&](uint32_t row_s, float (&d)[NMT][4]) {
      uint32_t b[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
      ldmatrix_x4_trans(row_s, b[0]);                      // y 0..31: k16 steps 0, 1
      if (NKS > 2) ldmatrix_x4_trans(row_s + 512u, b[1]);  // y 32..63: k16 steps 2, 3
      float dl[NMT][4];
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[mt][i] = dl[mt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          mma_bf16(d[mt], ah[mt][ks], b[ks >> 1][(ks & 1) * 2], b[ks >> 1][(ks & 1) * 2 + 1]);
          mma_bf16(dl[mt], al[mt][ks], b[ks >> 1][(ks & 1) * 2], b[ks >> 1][(ks & 1) * 2 + 1]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[mt][i] += dl[mt][i];
    };
    auto xreduce = [&](int r, const float (&d)[NMT][4]) {
      const int i0 = (int)(sx * (float)r);
      if (i0 != a) {  // i0 == a + 1 (scale <= 1): row a is complete
        if (a & 1) flush(a, accO); else flush(a, accE);
        a = i0;
      }
      const float wa = hat_weight_(r, a, sx), wb = hat_weight_(r, a + 1, sx);
      const float wE = (a & 1) ? wb : wa, wO = (a & 1) ? wa : wb;
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          accE[mt][i] = fmaf(wE, d[mt][i], accE[mt][i]);
          accO[mt][i] = fmaf(wO, d[mt][i], accO[mt][i]);
        }
    };
    for (int row0 = 0; row0 < u.nrows; row0 += kRows, ++cnt) {
      const int s = cnt % kStages;
      const int rows = min(kRows, u.nrows - row0);
      mbar_wait(&bar_full[warp][s], (cnt / kStages) & 1);
      const uint32_t stage_s = ring_s + (uint32_t)s * p.stage_bytes + (uint32_t)lane * 16u;
      const int rbase = u.r0 + row0;
      int rr = 0;
      for (; rr + 1 < rows; rr += 2) {  // two rows at a time: their MMA chains are independent
        float d0[NMT][4], d1[NMT][4];
        yreduce(stage_s + (uint32_t)rr * row_bytes, d0);
        yreduce(stage_s + (uint32_t)(rr + 1) * row_bytes, d1);
        xreduce(rbase + rr, d0);
        xreduce(rbase + rr + 1, d1);
      }
      if (rr < rows) {
        float d0[NMT][4];
        yreduce(stage_s + (uint32_t)rr * row_bytes, d0);
        xreduce(rbase + rr, d0);
      }
      // the stage has been consumed (every ldmatrix fed an mma issued above): refill it kStages chunks ahead
      __syncwarp();
      produce(s);
    }
    if (a & 1) { flush(a, accO); flush(a + 1, accE); } else { flush(a, accE); flush(a + 1, accO); }
  }
}

}  // namespace

// Called by rtp_upsample_bwd (fuse.cu).  Returns 1 when the launch was issued, 0 when the shape is not supported (the caller
// falls back to the CUDA-core kernels), < 0 on error.
int rtp_upsample_bwd_yx_mma(const rtp_p8& dout, const rtp_p8& t2, int C8, void* stream) {
  static const bool off = getenv("RTP_NO_UPBWD_MMA") != nullptr;  // A/B switch
  if (off) return 0;
  const int Y = dout.Y, Yl = t2.Y, X = dout.X, Xl = t2.X;
  if (Y % 16 != 0 || Y > 64 || Yl > 32 || Yl < 2 || Xl < 2 || X < 2 || Yl > Y || Xl > X) return 0;
  const float sx = (float)(Xl - 1) / (float)(X - 1);
  UBM k;
  k.in = P8(dout);
  k.out = P8(t2);
  k.C8 = C8;
  k.nks = Y / 16;
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  // a unit = one x segment of a (sample, chunk, z) plane, streamed kRows rows at a time.  A segment of txl outputs re-reads
  // ~1 / scale halo rows (efficiency txl / (txl + 1)); the warps take units round-robin, so the last round should be full
  // (efficiency rounds / ceil(rounds)): pick the segment count with the best product
  const int64_t planes = (int64_t)dout.N * C8 * dout.Z;
  const uint32_t row_bytes0 = (uint32_t)(Y + 2) * 16u;
  const int kStages = stages_of(Yl > 16 ? 2 : 1);
  const size_t smem0 = (size_t)kWarps * kStages * (((uint32_t)kRows * row_bytes0 + 127u) & ~127u) + 1024;
  int cta_per_sm = (int)((227 * 1024) / (smem0 + 1024));  // + 1 KB the system reserves per CTA
  const int reg_cap = Yl > 16 ? 3 : 4;                    // 161 / 96 registers x 128 threads
  if (cta_per_sm > reg_cap) cta_per_sm = reg_cap;
  if (cta_per_sm < 1) cta_per_sm = 1;
  const int64_t wtotal = (int64_t)nsm * cta_per_sm * kWarps;
  const int ntx_max = Xl / 2 > 1 ? Xl / 2 : 1;
  int best_ntx = 1;
  double best = -1.0;
  for (int ntx = 1; ntx <= ntx_max; ++ntx) {
    const int txl = ceil_div(Xl, ntx);
    const int ntx_eff = ceil_div(Xl, txl);
    const double rounds = (double)(ntx_eff * planes) / (double)wtotal;
    const double balance = rounds / (double)(int64_t)(rounds + 0.999999);
    const double score = balance * ((double)txl / (double)(txl + 1));
    if (score > best + 1e-9) { best = score; best_ntx = ntx_eff; }
  }
  k.txl = ceil_div(Xl, best_ntx);
  k.ntx = ceil_div(Xl, k.txl);
  const int64_t nunits = (int64_t)k.ntx * planes;
  if (nunits > 0x7fffffff) return 0;
  k.nunits = (int)nunits;
  const uint32_t row_bytes = (uint32_t)(Y + 2) * 16u;
  k.stage_bytes = ((uint32_t)kRows * row_bytes + 127u) & ~127u;
  // + 1 KB: ldmatrix always fetches 32 (or 64) y positions of a row; the k16 steps beyond Y are never multiplied
  const size_t smem = (size_t)kWarps * kStages * k.stage_bytes + 1024;
  (void)sx;
  const int nmt = Yl > 16 ? 2 : 1;
  using Kern = void (*)(const UBM);
  static const Kern kerns[2][4] = {
      {upsample_bwd_yx_mma_kernel<1, 1>, upsample_bwd_yx_mma_kernel<1, 2>, upsample_bwd_yx_mma_kernel<1, 3>, upsample_bwd_yx_mma_kernel<1, 4>},
      {upsample_bwd_yx_mma_kernel<2, 1>, upsample_bwd_yx_mma_kernel<2, 2>, upsample_bwd_yx_mma_kernel<2, 3>, upsample_bwd_yx_mma_kernel<2, 4>}};
  const Kern kern = kerns[nmt - 1][k.nks - 1];
  static size_t configured_dev[2][4][RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[nmt - 1][k.nks - 1][rtp_current_device()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_upsample_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    configured = smem;
  }
  const int want = ceil_div(k.nunits, kWarps);
  const int grid = want < cta_per_sm * nsm ? want : cta_per_sm * nsm;
  kern<<<grid, kWarps * 32, smem, (cudaStream_t)stream>>>(k);
  return 1;
}
