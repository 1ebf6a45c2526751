// wgrad_k3s1.cu — weight gradient of the 3x3x3 stride-1 convs with 32 input channels (the dominant shape) on tcgen05.
//
//   dW[kz][ky][kx][ci][co] = sum_{n,z,q} X[n, z+kz-1, q + (kx-1)*Yp + (ky-1)][ci] * dY[n, z, q][co]
//
// GEMM view: K = in-plane positions q (both operands MN-major straight out of the P8 layout).  A stacks the three
// X planes z-1, z, z+1 along M ((kz, ci) = 96 rows, padded to the UMMA M = 128), B is the dY plane (N = Cout padded
// to 16); the 9 in-plane taps are 9 accumulators [128 x NP] resident in TMEM, addressed — like in conv_k3s1.cu — by
// shifting the A descriptor's start address over ONE staged copy of the input rows.  Each persistent CTA walks
// (sample, 128-position tile) units over all z planes and writes one fp32 partial at the end; a second kernel sums
// the per-CTA partials in a fixed order (deterministic) into the reference's [Cout][Cin][3][3][3] layout.
//
// Roofline: tensor pipe; SMEM operand bandwidth (128 B/clk) caps M=128,N=32 at ~35 % and one of the four stacked
// plane slots is padding, so the ceiling is ~27 % of the bf16 peak (see DESIGN.md for the N=64 two-role variant).
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 final epilogue
constexpr int kMaxStages = 4;

struct WG3 {
  P8 x, dy;
  const bf16* zero_page;  // >= PW*16*4 bytes of zeros (stands in for the planes z = -1 and z = Z)
  int NP;                 // dY channels padded to 16
  int PW;                 // staged positions per plane: 128 + 2*Yp + 2
  int ntile, nunits, nstages;
  int valid_pos;          // X*Yp: positions past this in the last tile are skipped in whole k16 steps
  uint32_t xplane_bytes;  // 4 chunks * PW * 16
  uint32_t stage_bytes;   // 4 plane slots + dY tile
  float* partial;         // [gridDim.x][9][128][NP]
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_k3s1_kernel(const __grid_constant__ WG3 p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.nstages, Z = p.x.Z, Yp = p.x.Yp;
  const uint32_t dy_bytes = (uint32_t)(p.NP / 8) * 128 * 16;

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
  const bool has_work = (int)blockIdx.x < p.nunits;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const int tile = u % p.ntile, n = u / p.ntile;
        const int64_t qoff = ((int64_t)tile * 128 - 1) * 8;      // staged X rows start at q0 - Yp - 1
        const int64_t q0 = ((int64_t)Yp + (int64_t)tile * 128) * 8;
        const bf16* xn = p.x.ptr + (int64_t)n * p.x.n_stride + qoff;
        const bf16* dn = p.dy.ptr + (int64_t)n * p.dy.n_stride + q0;
        for (int z = 0; z < Z; ++z) {
          const int s = it % S;
          mbar_wait(&bar_empty[s], ((it / S) & 1) ^ 1);
          mbar_arrive_expect_tx(&bar_full[s], 3 * p.xplane_bytes + dy_bytes);
          uint8_t* dst = smem + (size_t)s * p.stage_bytes;
          for (int i = 0; i < 3; ++i) {
            const int zx = z + i - 1;
            const bool ok = zx >= 0 && zx < Z;
            for (int c = 0; c < 4; ++c) {
              const bf16* src = ok ? xn + (int64_t)c * p.x.c_stride + (int64_t)zx * p.x.plane_elems() : p.zero_page;
              bulk_g2s(dst + (size_t)(i * 4 + c) * p.PW * 16, src, p.PW * 16, &bar_full[s]);
            }
          }
          uint8_t* ddst = dst + 4 * p.xplane_bytes;
          for (int c = 0; c < p.NP / 8; ++c) {
            const bf16* src = c < p.dy.C8 ? dn + (int64_t)c * p.dy.c_stride + (int64_t)z * p.dy.plane_elems() : p.zero_page;
            bulk_g2s(ddst + (size_t)c * 2048, src, 2048, &bar_full[s]);
          }
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform control flow; only the MMA / commit instructions are predicated on the leader lane
      const bool leader = lane == 0;
      uint32_t it = 0;
      bool first = true;
      const uint32_t idesc = idesc_bf16(128, p.NP, 1, 1);
      const uint32_t a_hi = (uint32_t)p.PW | (1u << 14), b_hi = 128u | (1u << 14);
      const uint32_t smem0 = smem_u32(smem);
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const int tile = u % p.ntile;
        int nk16 = (p.valid_pos - tile * 128 + 15) / 16;  // whole 16-position K steps inside the plane
        nk16 = nk16 > 8 ? 8 : nk16;
        for (int z = 0; z < Z; ++z) {
          const int s = it % S;
          mbar_wait(&bar_full[s], (it / S) & 1);
          fence_after_sync();
          // descriptor halves (SWIZZLE_NONE, MN-major): lo = start>>4 | (LBO = 128 B)>>4 << 16 ; hi = SBO>>4 | version
          const uint32_t xbase = smem0 + (uint32_t)s * p.stage_bytes;
          const uint32_t a_lo = (8u << 16) + (xbase >> 4);
          const uint32_t b_lo = (8u << 16) + ((xbase + 4 * p.xplane_bytes) >> 4);
          if (elect_one()) {  // elect.sync => no per-MMA waterfall loop in SASS
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const uint32_t at = a_lo + (uint32_t)((t9 / 3) * Yp + (t9 % 3));
#pragma unroll
              for (int k16 = 0; k16 < 8; ++k16) {
                if (k16 < nk16)
                  mma_ss(tmem + t9 * p.NP, ((uint64_t)a_hi << 32) | (at + k16 * 16), ((uint64_t)b_hi << 32) | (b_lo + k16 * 16),
                         idesc, (first && k16 == 0) ? 0u : 1u);
              }
            }
            mma_commit(&bar_empty[s]);
          }
          __syncwarp();
          first = false;
          ++it;
        }
      }
      if (has_work && leader) mma_commit(&bar_done);
    }
  } else {
    // final epilogue: 9 accumulators -> fp32 partial of this CTA
    const int lane_q = warp & 3;
    const int r = lane_q * 32 + lane;
    if (has_work) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
    }
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16);
    for (int t9 = 0; t9 < 9; ++t9) {
      float* dst = p.partial + (((size_t)blockIdx.x * 9 + t9) * 128 + r) * p.NP;
      for (int c16 = 0; c16 * 16 < p.NP; ++c16) {
        uint32_t v[16];
        if (has_work) {
          tmem_ld16(trow + t9 * p.NP + c16 * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (r < 96) {
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

// partial[split][t9 = kx*3+ky][(kz*4 + c)*8 + ci8][n] -> dW[co][ci0 + ci][kz][ky][kx]
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
      const int t9 = kx * 3 + ky;
      const int m = (kz * 4 + (ci >> 3)) * 8 + (ci & 7);
      const float* src = partial + ((size_t)t9 * 128 + m) * NP + n0 + co;
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

int plan_stages(int NP, int Y, uint32_t& xplane_bytes, uint32_t& stage_bytes, int& PW) {
  PW = 128 + 2 * (Y + 2) + 2;
  xplane_bytes = 4u * PW * 16;
  stage_bytes = 4 * xplane_bytes + (uint32_t)(NP / 8) * 2048;
  int S = (int)((220 * 1024) / stage_bytes);
  return S > kMaxStages ? kMaxStages : S;
}

}  // namespace

extern "C" int64_t rtp_wgrad_k3s1_workspace_bytes(int32_t NP, int32_t nsm) { return (int64_t)nsm * 9 * 128 * NP * 4; }
extern "C" int64_t rtp_wgrad_k3s1_zero_bytes(int32_t Y) { return (int64_t)(128 + 2 * (Y + 2) + 2) * 16 + 2048; }

extern "C" int rtp_wgrad_k3s1_supported(int32_t Cin, int32_t NP, int32_t Z, int32_t X, int32_t Y) {
  if (Cin != 32 || NP % 16 != 0 || NP < 16 || 9 * NP > 512 || Y < 6) return 0;
  uint32_t a, b;
  int PW;
  return plan_stages(NP, Y, a, b, PW) >= 2 && Z >= 1 && X >= 1 ? 1 : 0;
}

extern "C" int rtp_wgrad_k3s1(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                              void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && zero_page && workspace && nsplit_out, "rtp_wgrad_k3s1: null argument");
  RTP_CHECK_ARG(x.N == dy.N && x.Z == dy.Z && x.X == dy.X && x.Y == dy.Y, "rtp_wgrad_k3s1: geometry mismatch");
  RTP_CHECK_ARG(x.C8 >= 4 && rtp_wgrad_k3s1_supported(32, NP, x.Z, x.X, x.Y), "rtp_wgrad_k3s1: unsupported shape NP=%d", NP);
  RTP_CHECK_ARG(x.c_stride == (int64_t)x.Z * (x.X + 2) * (x.Y + 2) * 8 && dy.c_stride == x.c_stride,
                "rtp_wgrad_k3s1: planes must be contiguous per channel chunk");
  WG3 k;
  k.x = P8(x); k.dy = P8(dy); k.zero_page = (const bf16*)zero_page; k.NP = NP;
  k.nstages = plan_stages(NP, x.Y, k.xplane_bytes, k.stage_bytes, k.PW);
  const int Yp = x.Y + 2;
  k.valid_pos = x.X * Yp;
  k.ntile = (k.valid_pos + 127) / 128;
  k.nunits = x.N * k.ntile;
  k.partial = workspace;
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  *nsplit_out = grid;
  const size_t smem = (size_t)k.nstages * k.stage_bytes;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_k3s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_wgrad_k3s1: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  wgrad_k3s1_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
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
