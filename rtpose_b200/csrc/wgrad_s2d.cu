// wgrad_s2d.cu — weight gradient of a 3x3x3 STRIDE-2 conv with 32 input channels whose (normalised) input is held as the
// space-to-depth view (include/rtpose_b200.h, rtp_gn_apply_s2d), plane-streaming on tcgen05.
//
// Tap k of one dimension reads parity 1 at offset -1 (k = 0), parity 0 at offset 0 (k = 1) or parity 1 at offset 0
// (k = 2), so with view chunk (pz, px, py, c) = ((pz<<2 | px<<1 | py) * 4 + c):
//
//   dW[co][ci][kz][ky][kx] = sum_{n,z,x,y} XS[n][(pz,px,py), ci][z + oz][x + ox][y + oy] * dY[n][co][z][x][y]
//
// with (p, o) = (1,-1) / (0,0) / (1,0) per dimension: 27 (parity group, offset) pairs out of 8 x 8.  The gather kernel
// (wgrad_generic over the view) re-reads X once per tap — 0.57 GB of L2 -> SM traffic per launch, 125 TFLOP/s.  Here, as in
// wgrad_k3s1.cu, GEMM K = in-plane positions and both operands are MN-major straight from the P8 layout, but:
//   A: the 32 chunks of ONE view plane zx of a 128-position tile (+ one row and one position of halo in front) are staged
//      once, slot order (py, pz, px, c).  An in-plane offset (ox, oy) in {-1, 0}^2 is a shift of the K start
//      ((ox*Yp + oy) positions = 16-byte units of the descriptor start).  M = 128 rows = the 16 chunks of the four parity
//      groups with the same py — no padding rows.
//   B: N = 2*NP = (jz, co): the dY planes zx (oz = 0) and zx+1 (oz = -1) from a ring of dY plane tiles (plane Z = zero page).
//   6 accumulators [128 x 2*NP] in TMEM: block py=0 needs oy = 0 only -> ox in {-1, 0}; block py=1 needs all four (ox, oy).
//      A group that does not need a shift (e.g. px = 0 with ox = -1) gets it computed anyway inside the M = 128 instruction;
//      the reduce kernel simply does not read those entries (27 of the 48 [32 x NP] blocks per jz-half are used).
// Per plane-step: 8 k16 x 6 MMAs (M128 N64 K16); X is fetched once with a 1.27x halo instead of 27x.
// Persistent CTAs walk (sample, tile) units over all z planes and write one fp32 partial; rtp_wgrad_s2d_reduce sums the
// partials in a fixed order (deterministic) into the reference's [Cout][Cin][3][3][3] layout.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 final epilogue
constexpr int kStages = 2;
constexpr int kAcc = 6;

struct WS2D {
  P8 xs, dy;
  const bf16* zero_page;  // >= NP/8 * 2048 bytes of zeros (the dY plane z = Z)
  int NP, C8in;           // dY channels padded to 16; input chunks per parity group (4)
  int ntile, nunits, valid_pos;
  int PW;                 // staged positions per chunk: 128 + Yp + 1
  int R;                  // dY ring slots (+1 mirror of slot 0 behind them)
  uint32_t chunk_bytes, stage_bytes, slot_bytes;
  float* partial;         // [grid][6][128][2*NP]
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_s2d_kernel(const __grid_constant__ WS2D p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = p.R, Z = p.xs.Z, Yp = p.xs.Yp, ZP = Z + 1;  // planes 0..Z of dY per unit (plane Z = zeros)
  uint8_t* ring = smem + (size_t)kStages * p.stage_bytes;
  const int nch = p.NP / 8;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const bool has_work = (int)blockIdx.x < p.nunits;

  if (warp == 0) {
    // ------------------------------------------------------------ producer: lane 0 runs the barrier protocol, all lanes copy
    uint32_t it = 0, gbase = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, gbase += ZP) {
      const int tile = u % p.ntile, n = u / p.ntile;
      const int64_t q0 = (int64_t)Yp + (int64_t)tile * 128;
      const bf16* xn = p.xs.ptr + (int64_t)n * p.xs.n_stride + (q0 - Yp - 1) * 8;
      const bf16* dn = p.dy.ptr + (int64_t)n * p.dy.n_stride + q0 * 8;
      for (int zx = 0; zx < Z; ++zx, ++it) {
        const int s = it % kStages;
        // dY planes fetched with this step: 0 and 1 at the start of a unit, zx+1 afterwards
        const int pz_lo = zx == 0 ? 0 : zx + 1, npl = zx == 0 ? 2 : 1;
        if (lane == 0) {
          mbar_wait(&bar_empty[s], ((it / kStages) & 1) ^ 1);
          uint32_t bytes = p.stage_bytes;
          for (int i = 0; i < npl; ++i) {
            const uint32_t g = gbase + pz_lo + i, sl = g % R;
            bytes += p.slot_bytes * (sl == 0 ? 2u : 1u);
            if (g >= (uint32_t)R) {
              // the slot's previous plane P was last read by the step zx = min(P, Z-1) of its unit: that step must have retired
              const uint32_t gp = g - R, kprev = gp / ZP, pzp = gp % ZP;
              const uint32_t t = kprev * Z + (pzp < (uint32_t)Z ? pzp : (uint32_t)(Z - 1));
              if (t + kStages > it) mbar_wait(&bar_empty[t % kStages], (t / kStages) & 1);
            }
          }
          mbar_arrive_expect_tx(&bar_full[s], bytes);
        }
        __syncwarp();
        uint8_t* xdst = smem + (size_t)s * p.stage_bytes;
        const bf16* xz = xn + (int64_t)zx * p.xs.plane_elems();
        const int nx = 8 * p.C8in;                 // 32 view chunks
        const int ncopy = nx + npl * 2 * nch;
        for (int i = lane; i < ncopy; i += 32) {
          if (i < nx) {
            // smem slot i = (py, pz, px, c)  <-  view chunk ((pz<<2 | px<<1 | py) * C8in + c)
            const int c = i % p.C8in, q = i / p.C8in, px = q & 1, pz = (q >> 1) & 1, py = q >> 2;
            const int vc = ((pz << 2) | (px << 1) | py) * p.C8in + c;
            bulk_g2s(xdst + (size_t)i * p.chunk_bytes, xz + (int64_t)vc * p.xs.c_stride, p.chunk_bytes, &bar_full[s]);
          } else {
            const int d = i - nx, pi = d / (2 * nch), rem = d - pi * 2 * nch, c = rem >> 1, mirror = rem & 1;
            const int pzz = pz_lo + pi;
            const uint32_t sl = (gbase + pzz) % R;
            if (mirror && sl != 0) continue;
            const bool real = pzz < Z && c < p.dy.C8;
            const bf16* src = real ? dn + (int64_t)c * p.dy.c_stride + (int64_t)pzz * p.dy.plane_elems() : p.zero_page;
            bulk_g2s(ring + (size_t)(mirror ? R : sl) * p.slot_bytes + (size_t)c * 2048, src, 2048, &bar_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control flow, elect.sync issue blocks)
    uint32_t it = 0, gbase = 0;
    bool first = true;
    const int N3 = 2 * p.NP;
    const uint32_t idesc = idesc_bf16(128, N3, 1, 1);
    const uint32_t a_hi = (p.chunk_bytes >> 4) | (1u << 14), b_hi = 128u | (1u << 14);  // SBO: chunk stride / 2048 B
    const uint32_t smem0 = smem_u32(smem), ring0 = smem_u32(ring);
    const uint32_t blk1 = (16u * p.chunk_bytes) >> 4;  // second M block (py = 1) in 16-byte units
    // K-start shifts (positions = 16-byte units): the stage begins at q0 - Yp - 1
    const uint32_t sh00 = (uint32_t)Yp + 1, sh0m = (uint32_t)Yp, shm0 = 1u, shmm = 0u;  // (ox, oy) = (0,0) (0,-1) (-1,0) (-1,-1)
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, gbase += ZP) {
      const int tile = u % p.ntile;
      int nk16 = (p.valid_pos - tile * 128 + 15) / 16;  // whole 16-position K steps inside the plane
      nk16 = nk16 > 8 ? 8 : nk16;
      for (int zx = 0; zx < Z; ++zx, ++it) {
        const int s = it % kStages;
        mbar_wait(&bar_full[s], (it / kStages) & 1);
        fence_after_sync();
        const uint32_t a_lo = (8u << 16) + ((smem0 + (uint32_t)s * p.stage_bytes) >> 4);
        const uint32_t b_lo = (8u << 16) + ((ring0 + ((gbase + zx) % R) * p.slot_bytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            if (k16 >= nk16) continue;
            const uint32_t acc = (first && k16 == 0) ? 0u : 1u;
            const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + k16 * 16);
            const uint32_t a0 = a_lo + k16 * 16, a1 = a0 + blk1;
            mma_ss(tmem + 0 * N3, ((uint64_t)a_hi << 32) | (a0 + shm0), bd, idesc, acc);  // py=0: ox = -1
            mma_ss(tmem + 1 * N3, ((uint64_t)a_hi << 32) | (a0 + sh00), bd, idesc, acc);  // py=0: ox =  0
            mma_ss(tmem + 2 * N3, ((uint64_t)a_hi << 32) | (a1 + shmm), bd, idesc, acc);  // py=1: ox = -1, oy = -1
            mma_ss(tmem + 3 * N3, ((uint64_t)a_hi << 32) | (a1 + shm0), bd, idesc, acc);  // py=1: ox = -1, oy =  0
            mma_ss(tmem + 4 * N3, ((uint64_t)a_hi << 32) | (a1 + sh0m), bd, idesc, acc);  // py=1: ox =  0, oy = -1
            mma_ss(tmem + 5 * N3, ((uint64_t)a_hi << 32) | (a1 + sh00), bd, idesc, acc);  // py=1: ox =  0, oy =  0
          }
          mma_commit(&bar_empty[s]);
        }
        __syncwarp();
        first = false;
      }
    }
    if (has_work && lane == 0) mma_commit(&bar_done);
  } else {
    // ------------------------------------------------------------ final epilogue: 6 accumulators -> fp32 partial of this CTA
    const int lane_q = warp & 3;
    const int r = lane_q * 32 + lane;
    const int N3 = 2 * p.NP;
    if (has_work) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
    }
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16);
    for (int a = 0; a < kAcc; ++a) {
      float* dst = p.partial + (((size_t)blockIdx.x * kAcc + a) * 128 + r) * N3;
      for (int c16 = 0; c16 * 16 < N3; ++c16) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(trow + a * N3 + c16 * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + c16 * 16 + i) =
              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// partial[split][acc][((pz*2+px)*4 + ci/8)*8 + ci%8][jz*NP + co]  ->  dW[co][ci][kz][ky][kx]
__global__ void __launch_bounds__(256) wgrad_s2d_reduce_kernel(const float* __restrict__ partial, int nsplit, int NP, int Cin,
                                                               float* __restrict__ dW, int co_n, int accumulate) {
  __shared__ float sh[8][33];
  const int total = 27 * Cin * co_n;
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int N3 = 2 * NP;
  const size_t sstride = (size_t)kAcc * 128 * N3;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + o;
    float acc = 0.f;
    int co = 0, ci = 0, tap = 0;
    if (i < total) {
      co = i % co_n;
      const int r = i / co_n;
      ci = r % Cin;
      tap = r / Cin;  // (kz*3 + ky)*3 + kx
      const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
      // tap k -> (parity, offset): 0 -> (1,-1), 1 -> (0,0), 2 -> (1,0)
      const int pz = kz != 1, px = kx != 1, py = ky != 1;
      const int jz = kz == 0, oxm = kx == 0, oym = ky == 0;  // 1 when the offset is -1
      const int a = py == 0 ? (oxm ? 0 : 1) : 2 + (oxm ? 0 : 2) + (oym ? 0 : 1);
      const int m = ((pz * 2 + px) * (Cin / 8) + (ci >> 3)) * 8 + (ci & 7);
      const float* src = partial + ((size_t)a * 128 + m) * N3 + jz * NP + co;
      for (int s = sl; s < nsplit; s += 8) acc += src[s * sstride];
    }
    sh[sl][o] = acc;
    __syncthreads();
    if (sl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += sh[l][o];
      float* d = dW + ((int64_t)co * Cin + ci) * 27 + tap;
      *d = accumulate ? *d + t : t;
    }
    __syncthreads();
  }
}

bool plan(int NP, int Y, int& R, int& PW, uint32_t& chunk_bytes, uint32_t& stage_bytes, uint32_t& slot_bytes, size_t& smem) {
  const int Yp = Y + 2;
  PW = 128 + Yp + 1;
  chunk_bytes = (uint32_t)PW * 16u;
  stage_bytes = 32u * chunk_bytes;
  slot_bytes = (uint32_t)(NP / 8) * 2048u;
  const size_t budget = 220 * 1024;
  if ((size_t)kStages * stage_bytes + 4 * (size_t)slot_bytes > budget) return false;
  R = (int)((budget - (size_t)kStages * stage_bytes) / slot_bytes) - 1;
  if (R > 12) R = 12;
  if (R < 3) return false;
  smem = (size_t)kStages * stage_bytes + (size_t)(R + 1) * slot_bytes;
  return true;
}

}  // namespace

extern "C" int rtp_wgrad_s2d_supported(int32_t Cin, int32_t NP, int32_t Z, int32_t X, int32_t Y) {
  if (Cin != 32 || NP % 16 != 0 || NP < 16 || kAcc * 2 * NP > 512 || Y < 6 || Z < 1 || X < 1) return 0;
  int R, PW;
  uint32_t cb, sb, slb;
  size_t smem;
  return plan(NP, Y, R, PW, cb, sb, slb, smem) ? 1 : 0;
}
extern "C" int64_t rtp_wgrad_s2d_workspace_bytes(int32_t NP, int32_t nsm) { return (int64_t)nsm * kAcc * 128 * 2 * NP * 4; }

extern "C" int rtp_wgrad_s2d(rtp_p8 xs, rtp_p8 dy, int32_t Cin, int32_t NP, const void* zero_page, float* workspace,
                             int32_t* nsplit_out, void* stream) {
  RTP_CHECK_ARG(xs.ptr && dy.ptr && zero_page && workspace && nsplit_out, "rtp_wgrad_s2d: null argument");
  RTP_CHECK_ARG(xs.N == dy.N && xs.Z == dy.Z && xs.X == dy.X && xs.Y == dy.Y, "rtp_wgrad_s2d: geometry mismatch");
  RTP_CHECK_ARG(xs.C8 == 8 * (Cin / 8) && rtp_wgrad_s2d_supported(Cin, NP, xs.Z, xs.X, xs.Y),
                "rtp_wgrad_s2d: unsupported shape Cin=%d NP=%d (view chunks %d)", Cin, NP, xs.C8);
  RTP_CHECK_ARG(xs.c_stride == (int64_t)xs.Z * (xs.X + 2) * (xs.Y + 2) * 8 && dy.c_stride == xs.c_stride,
                "rtp_wgrad_s2d: planes must be contiguous per channel chunk");
  WS2D k;
  size_t smem;
  k.xs = P8(xs); k.dy = P8(dy); k.zero_page = (const bf16*)zero_page; k.NP = NP; k.C8in = Cin / 8;
  plan(NP, xs.Y, k.R, k.PW, k.chunk_bytes, k.stage_bytes, k.slot_bytes, smem);
  const int Yp = xs.Y + 2;
  k.valid_pos = xs.X * Yp;
  k.ntile = (k.valid_pos + 127) / 128;
  k.nunits = xs.N * k.ntile;
  k.partial = workspace;
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
    cudaError_t e = cudaFuncSetAttribute(wgrad_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_wgrad_s2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  wgrad_s2d_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_wgrad_s2d_reduce(const float* workspace, int32_t nsplit, int32_t Cin, int32_t NP, float* dW, int32_t co_n,
                                    int32_t accumulate, void* stream) {
  RTP_CHECK_ARG(workspace && dW && nsplit >= 1 && co_n >= 1 && co_n <= NP && Cin == 32, "rtp_wgrad_s2d_reduce: bad args");
  const int total = 27 * Cin * co_n;
  wgrad_s2d_reduce_kernel<<<ceil_div(total, 32), 256, 0, (cudaStream_t)stream>>>(workspace, nsplit, NP, Cin, dW, co_n, accumulate);
  RTP_LAUNCH_CHECK();
}
