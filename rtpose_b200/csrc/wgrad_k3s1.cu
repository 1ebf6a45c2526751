// wgrad_k3s1.cu — weight gradient of the 3x3x3 stride-1 convs with 32 input channels (the dominant shape) on tcgen05.
//
//   dW[kz][ky][kx][ci][co] = sum_{n,z,q} X[n, z+kz-1, q + (kx-1)*Yp + (ky-1)][ci] * dY[n, z, q][co]
//
// GEMM view: K = in-plane positions q, both operands MN-major straight out of the P8 layout.  Each step consumes ONE
// X plane zx of a 128-position tile:
//   A (M = 16 groups of 8 rows, group 4c+kx = channel chunk c of the X tile shifted by (kx-1) rows; groups 4c+3 are
//     padding): the +-1 position shift (ky) is a 16-byte shift of the descriptor start.  When a row is at least 65
//     positions wide ("span mode") the four groups of a chunk are ONE staged span of 4 rows read at a group stride of
//     one row (SBO = Yp*16 B), so X is fetched once, in 4 bulk copies per step; narrower planes stage three shifted
//     copies per chunk (12 copies of 130 positions).  Few, large copies matter: a bulk copy costs ~30 ns + bytes at
//     ~68 GB/s per SM (tools/bulk_probe.cu).
//   B (N = (jz, co) = 3*NP columns): the dY planes zx-1, zx, zx+1 (kz = 2-jz), which sit in consecutive slots of a
//     ring of dY plane tiles.  Every dY plane is fetched once per tile; the first two ring slots are mirrored behind
//     the last one so a 3-plane window never wraps.  Planes -1 and Z are copies of a zero page.
// => 3 accumulators [128 x 3*NP] (one per ky) live in TMEM for the whole kernel; an MMA is M128 x N96 x K16
// (56 issue cycles by the measured operand-bandwidth law, vs 3 x 45 for the N=32 form this replaces).  Each
// persistent CTA walks (sample, tile) units over all z planes and writes one fp32 partial at the end; a second kernel
// sums the per-CTA partials in a fixed order (deterministic) into the reference's [Cout][Cin][3][3][3] layout.
//
// Roofline: tensor pipe.  96 of the 128 M rows are useful, so the ceiling is 75 % of the N=96 issue rate.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 final epilogue
constexpr int kMaxStages = 4;
constexpr int kXW = 130;                       // staged positions per X copy: 128 + the two ky halo positions
constexpr uint32_t kXCopyStage = 16u * kXW * 16u;  // copy mode: 16 group slots (4 of them unused)

struct WG3 {
  P8 x, dy;
  const bf16* zero_page;  // >= 2048 bytes of zeros (stands in for the dY planes z = -1 and z = Z)
  int NP;                 // dY channels padded to 16
  int ntile, nunits;
  const int* unit_list;   // optional: the unit ids to process (rtp_active_units) and their count, else all nunits
  const int* unit_count;
  int nstages;            // X stages (one step each)
  int R;                  // dY ring slots (+2 mirror slots behind them)
  int valid_pos;          // X*Yp: positions past this in the last tile are skipped in whole k16 steps
  uint32_t slot_bytes;    // NP/8 chunks x 128 positions x 16 B
  float* partial;         // [gridDim.x][3 ky][128][3*NP]
  int span;               // 1: span mode (see header)
  uint32_t xstage_bytes;  // 16 groups x group stride
  uint32_t a_sbo;         // A group stride in 16-byte units: Yp (span) or kXW (copies)
  int dbg;                // tools only: 1 = no X copies, 2 = no MMAs, 4 = no dY copies (results are garbage)
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_k3s1_kernel(const __grid_constant__ WG3 p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.nstages, R = p.R, Z = p.x.Z, Yp = p.x.Yp, ZP = Z + 2;
  uint8_t* ring = smem + (size_t)S * p.xstage_bytes;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const int nloop = p.unit_list ? __ldg(p.unit_count) : p.nunits;
  const bool has_work = (int)blockIdx.x < nloop;

  if (warp == 0) {
    // producer warp: lane 0 runs the barrier protocol, then all lanes issue the step's bulk copies in parallel
    uint32_t it = 0, gbase = 0;  // CTA-local step counter / plane counter at the start of the unit
    const int nch = p.NP / 8;
    for (int ku = blockIdx.x; ku < nloop; ku += gridDim.x, gbase += ZP) {
      const int u = p.unit_list ? p.unit_list[ku] : ku;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int64_t q0 = (int64_t)Yp + (int64_t)tile * 128;
      const bf16* xn = p.x.ptr + (int64_t)n * p.x.n_stride + (q0 - 1 - Yp) * 8;
      const bf16* dn = p.dy.ptr + (int64_t)n * p.dy.n_stride + q0 * 8;
      for (int zx = 0; zx < Z; ++zx, ++it) {
        const int s = it % S;
        // dY planes fetched with this step: -1, 0, 1 at the start of a unit, zx+1 afterwards
        const int pz_lo = zx == 0 ? -1 : zx + 1, npl = zx == 0 ? 3 : 1;
        if (lane == 0) {
          mbar_wait(&bar_empty[s], ((it / S) & 1) ^ 1);
          uint32_t bytes = (p.dbg & 1) ? 0u : (p.span ? p.xstage_bytes : 12u * kXW * 16u);
          for (int i = 0; i < npl; ++i) {
            const uint32_t g = gbase + pz_lo + i + 1, sl = g % R;
            if (!(p.dbg & 4)) bytes += p.slot_bytes * (sl < 2 ? 2u : 1u);
            if (g >= (uint32_t)R) {
              // the slot's previous plane was last read by step t of this CTA; make sure that step has retired
              const uint32_t gp = g - R, kprev = gp / ZP;
              const int pzp = (int)(gp % ZP) - 1;
              const uint32_t t = kprev * Z + (uint32_t)(pzp + 1 < Z - 1 ? pzp + 1 : Z - 1);
              if (t + S > it) mbar_wait(&bar_empty[t % S], (t / S) & 1);
            }
          }
          mbar_arrive_expect_tx(&bar_full[s], bytes);
        }
        __syncwarp();
        uint8_t* xdst = smem + (size_t)s * p.xstage_bytes;
        const bf16* xz = xn + (int64_t)zx * p.x.plane_elems();
        const int nx = p.span ? 4 : 12;
        const int ncopy = nx + npl * 2 * nch;
        for (int i = lane; i < ncopy; i += 32) {
          if (p.dbg & (i < nx ? 1 : 4)) continue;
          if (i < nx) {
            if (p.span) {  // chunk i: 4 rows starting one row above the tile
              bulk_g2s(xdst + (size_t)i * (p.xstage_bytes >> 2), xz + (int64_t)i * p.x.c_stride, p.xstage_bytes >> 2, &bar_full[s]);
            } else {       // chunk c shifted by j rows -> group slot 4c + j
              const int j = i >> 2, c = i & 3;
              bulk_g2s(xdst + (size_t)(4 * c + j) * (kXW * 16), xz + (int64_t)c * p.x.c_stride + (int64_t)j * Yp * 8, kXW * 16,
                       &bar_full[s]);
            }
          } else {
            const int d = i - nx, pi = d / (2 * nch), rem = d - pi * 2 * nch, c = rem >> 1, mirror = rem & 1;
            const int pz = pz_lo + pi;
            const uint32_t sl = (gbase + pz + 1) % R;
            if (mirror && sl >= 2) continue;
            const bool real = pz >= 0 && pz < Z && c < p.dy.C8;
            const bf16* src = real ? dn + (int64_t)c * p.dy.c_stride + (int64_t)pz * p.dy.plane_elems() : p.zero_page;
            bulk_g2s(ring + (size_t)(mirror ? R + sl : sl) * p.slot_bytes + (size_t)c * 2048, src, 2048, &bar_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform control flow; the MMA / commit instructions sit in elect.sync blocks
    uint32_t it = 0, gbase = 0;
    bool first = true;
    const int N3 = 3 * p.NP;
    const uint32_t idesc = idesc_bf16(128, N3, 1, 1);
    const uint32_t a_hi = p.a_sbo | (1u << 14), b_hi = 128u | (1u << 14);
    const uint32_t smem0 = smem_u32(smem), ring0 = smem_u32(ring);
    for (int ku = blockIdx.x; ku < nloop; ku += gridDim.x, gbase += ZP) {
      const int u = p.unit_list ? p.unit_list[ku] : ku;
      const int tile = u % p.ntile;
      int nk16 = (p.valid_pos - tile * 128 + 15) / 16;  // whole 16-position K steps inside the plane
      nk16 = nk16 > 8 ? 8 : nk16;
      for (int zx = 0; zx < Z; ++zx, ++it) {
        const int s = it % S;
        mbar_wait(&bar_full[s], (it / S) & 1);
        fence_after_sync();
        // descriptor halves (SWIZZLE_NONE, MN-major): lo = start>>4 | (LBO = 128 B)>>4 << 16 ; hi = SBO>>4 | version
        const uint32_t a_lo = (8u << 16) + ((smem0 + (uint32_t)s * p.xstage_bytes) >> 4);
        const uint32_t b_lo = (8u << 16) + ((ring0 + ((gbase + zx) % R) * p.slot_bytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int k16 = 0; k16 < 8; ++k16) {
              if (k16 < nk16 && !((p.dbg & 2) && it > 0))
                mma_ss(tmem + ky * N3, ((uint64_t)a_hi << 32) | (a_lo + ky + k16 * 16), ((uint64_t)b_hi << 32) | (b_lo + k16 * 16),
                       idesc, (first && k16 == 0) ? 0u : 1u);
            }
          }
          mma_commit(&bar_empty[s]);
        }
        __syncwarp();
        first = false;
      }
    }
    if (has_work && lane == 0) mma_commit(&bar_done);
  } else {
    // final epilogue: 3 accumulators -> fp32 partial of this CTA
    const int lane_q = warp & 3;
    const int r = lane_q * 32 + lane;
    const int N3 = 3 * p.NP;
    if (has_work) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
    }
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16);
    for (int ky = 0; ky < 3; ++ky) {
      float* dst = p.partial + (((size_t)blockIdx.x * 3 + ky) * 128 + r) * N3;
      for (int c16 = 0; c16 * 16 < N3; ++c16) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(trow + ky * N3 + c16 * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (((r >> 3) & 3) != 3) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + c16 * 16 + i) =
                make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// partial[split][ky][((ci/8)*4 + kx)*8 + ci%8][(2-kz)*NP + n] -> dW[co][ci0 + ci][kz][ky][kx]
__global__ void __launch_bounds__(256) wgrad_k3s1_reduce_kernel(const float* __restrict__ partial, int nsplit, int NP,
                                                                float* __restrict__ dW, int Cin_total, int co_n, int n0,
                                                                int ci0, int accumulate) {
  // 32 outputs x 8 split lanes per block, fixed-order combine (deterministic)
  __shared__ float sh[8][33];
  const int total = 27 * 32 * co_n;
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const size_t sstride = (size_t)9 * 128 * NP;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + o;
    float acc = 0.f;
    int co = 0, ci = 0, tap = 0;
    if (i < total) {
      co = i % co_n;
      const int r = i / co_n;
      ci = r % 32;
      tap = r / 32;  // (kz*3 + ky)*3 + kx
      const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
      const float* src = partial + ((size_t)ky * 128 + ((ci >> 3) * 4 + kx) * 8 + (ci & 7)) * (3 * NP) + (2 - kz) * NP + n0 + co;
      for (int s = sl; s < nsplit; s += 8) acc += src[s * sstride];
    }
    sh[sl][o] = acc;
    __syncthreads();
    if (sl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += sh[l][o];
      float* d = dW + ((int64_t)co * Cin_total + ci0 + ci) * 27 + tap;
      *d = accumulate ? *d + t : t;
    }
    __syncthreads();
  }
}

// X stages + dY ring (R slots + 2 mirrors); returns false if even the smallest configuration does not fit
bool plan(int NP, int Y, int& S, int& R, int& span, uint32_t& xstage, uint32_t& slot_bytes, size_t& smem) {
  const int Yp = Y + 2;
  span = (Yp >= 65 && Yp <= kXW) ? 1 : 0;  // 130 + 2*Yp staged positions must fit the 4-row span
  xstage = span ? 16u * Yp * 16u : kXCopyStage;
  slot_bytes = (uint32_t)(NP / 8) * 2048;
  const size_t budget = 220 * 1024;
  for (S = kMaxStages; S >= 2; --S) {
    R = (int)((budget - (size_t)S * xstage) / slot_bytes) - 2;
    if (R > 14) R = 14;
    if (R >= S + 2) break;
  }
  if (S < 2) {
    S = 2;
    R = (int)((budget - 2 * (size_t)xstage) / slot_bytes) - 2;
    if (R < 3) return false;
  }
  smem = (size_t)S * xstage + (size_t)(R + 2) * slot_bytes;
  return true;
}

}  // namespace

extern "C" { int rtp_wgrad_k3s1_dbg = 0; }  // tools/dbg_wgrad.py only; not part of the ABI

extern "C" int64_t rtp_wgrad_k3s1_workspace_bytes(int32_t NP, int32_t nsm) { return (int64_t)nsm * 9 * 128 * NP * 4; }
extern "C" int64_t rtp_wgrad_k3s1_zero_bytes(int32_t Y) { return (int64_t)(128 + 2 * (Y + 2) + 2) * 16 + 2048; }

extern "C" int rtp_wgrad_k3s1_supported(int32_t Cin, int32_t NP, int32_t Z, int32_t X, int32_t Y) {
  if (Cin != 32 || NP % 16 != 0 || NP < 16 || 9 * NP > 512 || Y < 6) return 0;
  int S, R, span;
  uint32_t xs, sb;
  size_t smem;
  return plan(NP, Y, S, R, span, xs, sb, smem) && Z >= 1 && X >= 1 ? 1 : 0;
}

static int wgrad_k3s1_launch(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                             const int32_t* unit_list, const int32_t* unit_count, void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && zero_page && workspace && nsplit_out, "rtp_wgrad_k3s1: null argument");
  RTP_CHECK_ARG(x.N == dy.N && x.Z == dy.Z && x.X == dy.X && x.Y == dy.Y, "rtp_wgrad_k3s1: geometry mismatch");
  RTP_CHECK_ARG(x.C8 >= 4 && rtp_wgrad_k3s1_supported(32, NP, x.Z, x.X, x.Y), "rtp_wgrad_k3s1: unsupported shape NP=%d", NP);
  RTP_CHECK_ARG(x.c_stride == (int64_t)x.Z * (x.X + 2) * (x.Y + 2) * 8 && dy.c_stride == x.c_stride,
                "rtp_wgrad_k3s1: planes must be contiguous per channel chunk");
  WG3 k;
  size_t smem;
  k.x = P8(x); k.dy = P8(dy); k.zero_page = (const bf16*)zero_page; k.NP = NP;
  plan(NP, x.Y, k.nstages, k.R, k.span, k.xstage_bytes, k.slot_bytes, smem);
  const int Yp = x.Y + 2;
  k.a_sbo = k.span ? (uint32_t)Yp : (uint32_t)kXW;
  k.valid_pos = x.X * Yp;
  k.ntile = (k.valid_pos + 127) / 128;
  k.nunits = x.N * k.ntile;
  k.partial = workspace;
  k.unit_list = unit_list;
  k.unit_count = unit_count;
  k.dbg = rtp_wgrad_k3s1_dbg;
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  *nsplit_out = grid;
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_k3s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_wgrad_k3s1: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  wgrad_k3s1_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_wgrad_k3s1(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                              void* stream) {
  return wgrad_k3s1_launch(x, dy, NP, zero_page, workspace, nsplit_out, nullptr, nullptr, stream);
}
extern "C" int rtp_wgrad_k3s1_units(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                                    const int32_t* unit_list, const int32_t* unit_count, void* stream) {
  RTP_CHECK_ARG(unit_list && unit_count, "rtp_wgrad_k3s1_units: null unit list");
  return wgrad_k3s1_launch(x, dy, NP, zero_page, workspace, nsplit_out, unit_list, unit_count, stream);
}

extern "C" int rtp_wgrad_k3s1_reduce(const float* workspace, int32_t nsplit, int32_t NP, float* dW, int32_t Cin_total,
                                     int32_t co_n, int32_t n0, int32_t ci0, int32_t accumulate, void* stream) {
  RTP_CHECK_ARG(workspace && dW && nsplit >= 1 && co_n >= 1 && n0 >= 0 && n0 + co_n <= NP && ci0 >= 0 && ci0 + 32 <= Cin_total,
                "rtp_wgrad_k3s1_reduce: bad args");
  const int total = 27 * 32 * co_n;
  wgrad_k3s1_reduce_kernel<<<ceil_div(total, 32), 256, 0, (cudaStream_t)stream>>>(workspace, nsplit, NP, dW, Cin_total, co_n,
                                                                                  n0, ci0, accumulate);
  RTP_LAUNCH_CHECK();
}

int rtp_wgrad_k3s1_set_carveout(int pct) {  // see rtp_set_shared_carveout (layout.cu)
  return (int)cudaFuncSetAttribute((const void*)wgrad_k3s1_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
