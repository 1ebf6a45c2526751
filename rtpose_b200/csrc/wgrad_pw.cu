// wgrad_pw.cu — weight gradient of a 1x1x1 conv as a streaming GEMM over P8 positions on tcgen05.
//
//   dW[co][ci0 + ci] (=|+=) sum_{n, pos} dY[n][co][pos] * X[n][ci][pos]
//
// The gather kernel (wgrad_generic, one tap) fills 4 of its 16 (tap, chunk) pair slots for a 32-channel X — 75 % of every
// MMA is padding — and gathers 16 bytes at a time: 0.30 ms for the 32 -> 128 final conv at full resolution, whose operands
// (168 MB + 671 MB) HBM can deliver in 0.13 ms.  Here GEMM M = dY channels (128 = 16 chunks: no padding for the 128-channel
// final conv), N = X channels, K = positions, both operands MN-major straight from the P8 layout (channel chunk = 8 rows,
// 16 bytes per position), fetched by 1-D bulk copies of 128 consecutive positions per chunk.  Pads are zero in both
// tensors, so the kernel simply streams every padded position of a (sample, chunk) volume; the last (< 16) positions of a
// volume lie in the final pad row and are skipped.  Persistent CTAs, one accumulator [128 x N] in TMEM, one fp32 partial per
// CTA, fixed-order reduce (deterministic).
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 final epilogue
constexpr int kMaxStages = 4;

struct WPW {
  P8 x, dy;
  const bf16* zero_page;  // >= 2048 bytes of zeros (dY chunks beyond its channel count)
  const bf16* ones_page;  // 128 positions x 8 channels with channel 0 = 1 (the bias-gradient column), or nullptr
  int bias_chunk;         // N chunk fed from ones_page (column 8 * bias_chunk of the result = sum over positions of dY), -1: none
  int NX;                 // X channels (+ the bias chunk) padded to 16 (GEMM N)
  int P16;                // whole 16-position groups per (sample, chunk) volume
  int ntile, nunits, nstages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  float* partial;         // [grid][128][NX]
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_pw_kernel(const __grid_constant__ WPW p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.nstages, nxc = p.NX / 8;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const bool has_work = (int)blockIdx.x < p.nunits;

  if (warp == 0) {
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int tile = u % p.ntile, n = u / p.ntile;
      const int s = it % S;
      if (lane == 0) {
        mbar_wait(&bar_empty[s], ((it / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], p.stage_bytes);
      }
      __syncwarp();
      uint8_t* dst = smem + (size_t)s * p.stage_bytes;
      const int64_t q0 = (int64_t)tile * 128 * 8;  // element offset of the tile inside a chunk volume
      for (int i = lane; i < 16 + nxc; i += 32) {
        if (i < 16) {
          const bf16* src = i < p.dy.C8 ? p.dy.ptr + (int64_t)n * p.dy.n_stride + (int64_t)i * p.dy.c_stride + q0 : p.zero_page;
          bulk_g2s(dst + (size_t)i * 2048, src, 2048, &bar_full[s]);
        } else {
          const int c = i - 16;
          const bf16* src = c < p.x.C8 ? p.x.ptr + (int64_t)n * p.x.n_stride + (int64_t)c * p.x.c_stride + q0
                                       : (c == p.bias_chunk ? p.ones_page : p.zero_page);
          bulk_g2s(dst + p.a_bytes + (size_t)c * 2048, src, 2048, &bar_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0;
    bool first = true;
    const uint32_t idesc = idesc_bf16(128, p.NX, 1, 1);
    const uint32_t hi = 128u | (1u << 14);  // SBO = 2048 B between channel chunks
    const uint32_t smem0 = smem_u32(smem);
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int tile = u % p.ntile;
      int nk16 = p.P16 - tile * 8;  // whole 16-position K steps of this tile inside the volume
      nk16 = nk16 > 8 ? 8 : nk16;
      const int s = it % S;
      mbar_wait(&bar_full[s], (it / S) & 1);
      fence_after_sync();
      const uint32_t a_lo = (8u << 16) + ((smem0 + (uint32_t)s * p.stage_bytes) >> 4);
      const uint32_t b_lo = a_lo + (p.a_bytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          if (k16 < nk16)
            mma_ss(tmem, ((uint64_t)hi << 32) | (a_lo + k16 * 16), ((uint64_t)hi << 32) | (b_lo + k16 * 16), idesc,
                   (first && k16 == 0) ? 0u : 1u);
        }
        mma_commit(&bar_empty[s]);
      }
      __syncwarp();
      if (nk16 > 0) first = false;
    }
    if (has_work && lane == 0) mma_commit(&bar_done);
  } else {
    const int lane_q = warp & 3;
    const int r = lane_q * 32 + lane;
    if (has_work) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
    }
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16);
    float* dst = p.partial + ((size_t)blockIdx.x * 128 + r) * p.NX;
    for (int c16 = 0; c16 * 16 < p.NX; ++c16) {
      uint32_t v[16];
      if (has_work) {
        tmem_ld16(trow + c16 * 16, v);
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
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// partial[split][co][ci] -> dW[co][ci0 + ci]
// (bias_col >= 0: item co_n * ci_n + co is dbias[co] = column bias_col of row co, the product with the ones channel)
__global__ void __launch_bounds__(256) wgrad_pw_reduce_kernel(const float* __restrict__ partial, int nsplit, int NX, float* __restrict__ dW,
                                                              int Cin_total, int co_n, int ci0, int ci_n, int accumulate,
                                                              float* __restrict__ dbias, int bias_col, int accumulate_bias) {
  __shared__ float sh[8][33];
  const int nw = co_n * ci_n;
  const int total = nw + (bias_col >= 0 ? co_n : 0);
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const size_t sstride = (size_t)128 * NX;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + o;
    float acc = 0.f;
    int co = 0, ci = 0;
    if (i < total) {
      if (i < nw) {
        ci = i % ci_n;
        co = i / ci_n;
      } else {
        ci = bias_col;
        co = i - nw;
      }
      const float* src = partial + (size_t)co * NX + ci;
      for (int s = sl; s < nsplit; s += 8) acc += src[s * sstride];
    }
    sh[sl][o] = acc;
    __syncthreads();
    if (sl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += sh[l][o];
      if (i < nw) {
        float* d = dW + (int64_t)co * Cin_total + ci0 + ci;
        *d = accumulate ? *d + t : t;
      } else {
        dbias[co] = accumulate_bias ? dbias[co] + t : t;
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int rtp_wgrad_pw_supported(int32_t Cin, int32_t Cout, int32_t Z, int32_t X, int32_t Y) {
  // M = 128 dY-channel rows; N = X channels (16..256); the final pad row (Y + 2 positions) must cover the skipped tail
  return (Cin >= 8 && Cin <= 256 && Cout >= 1 && Cout <= 128 && Y + 2 >= 16 && Z >= 1 && X >= 1) ? 1 : 0;
}
static int nx_of(int Cin, int with_bias) { return ((Cin + 7) / 8 * 8 + (with_bias ? 8 : 0) + 15) / 16 * 16; }
extern "C" int64_t rtp_wgrad_pw_workspace_bytes(int32_t Cin, int32_t nsm) { return (int64_t)nsm * 128 * nx_of(Cin, 0) * 4; }
extern "C" int64_t rtp_wgrad_pw_bias_workspace_bytes(int32_t Cin, int32_t nsm) { return (int64_t)nsm * 128 * nx_of(Cin, 1) * 4; }

static int wgrad_pw_launch(rtp_p8 x, rtp_p8 dy, int32_t Cin, const void* zero_page, const void* ones_page, float* workspace,
                           int32_t* nsplit_out, void* stream) {
  RTP_CHECK_ARG(x.ptr && dy.ptr && zero_page && workspace && nsplit_out, "rtp_wgrad_pw: null argument");
  RTP_CHECK_ARG(x.N == dy.N && x.Z == dy.Z && x.X == dy.X && x.Y == dy.Y, "rtp_wgrad_pw: geometry mismatch");
  RTP_CHECK_ARG(Cin <= x.C8 * 8 && rtp_wgrad_pw_supported(Cin, dy.C8 * 8 > 128 ? 129 : 128, x.Z, x.X, x.Y) && dy.C8 <= 16,
                "rtp_wgrad_pw: unsupported shape Cin=%d dY chunks=%d", Cin, dy.C8);
  const int64_t vol = (int64_t)x.Z * (x.X + 2) * (x.Y + 2);
  RTP_CHECK_ARG(x.c_stride == vol * 8 && dy.c_stride == vol * 8, "rtp_wgrad_pw: planes must be contiguous per channel chunk");
  WPW k;
  k.x = P8(x); k.dy = P8(dy); k.zero_page = (const bf16*)zero_page;
  k.x.C8 = (Cin + 7) / 8;
  k.ones_page = (const bf16*)ones_page;
  k.bias_chunk = ones_page ? k.x.C8 : -1;
  k.NX = nx_of(Cin, ones_page != nullptr);
  RTP_CHECK_ARG(k.NX <= 256, "rtp_wgrad_pw: too many X channels");
  k.P16 = (int)(vol / 16);
  k.ntile = (k.P16 + 7) / 8;
  k.nunits = x.N * k.ntile;
  k.a_bytes = 16u * 2048u;
  k.b_bytes = (uint32_t)(k.NX / 8) * 2048u;
  k.stage_bytes = k.a_bytes + k.b_bytes;
  k.nstages = (int)((200 * 1024) / k.stage_bytes);
  if (k.nstages > kMaxStages) k.nstages = kMaxStages;
  RTP_CHECK_ARG(k.nstages >= 2, "rtp_wgrad_pw: stage too large");
  k.partial = workspace;
  const size_t smem = (size_t)k.nstages * k.stage_bytes;
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
    cudaError_t e = cudaFuncSetAttribute(wgrad_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_wgrad_pw: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  wgrad_pw_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_wgrad_pw(rtp_p8 x, rtp_p8 dy, int32_t Cin, const void* zero_page, float* workspace, int32_t* nsplit_out,
                            void* stream) {
  return wgrad_pw_launch(x, dy, Cin, zero_page, nullptr, workspace, nsplit_out, stream);
}
// The same launch with one more N chunk fed from `ones_page` (128 positions x 8 channels, channel 0 = 1): the product with it
// is sum over positions of dY = the conv's bias gradient, taken out by rtp_wgrad_pw_bias_reduce — dY is not read again for it.
extern "C" int rtp_wgrad_pw_bias(rtp_p8 x, rtp_p8 dy, int32_t Cin, const void* zero_page, const void* ones_page, float* workspace,
                                 int32_t* nsplit_out, void* stream) {
  RTP_CHECK_ARG(ones_page, "rtp_wgrad_pw_bias: null ones page");
  return wgrad_pw_launch(x, dy, Cin, zero_page, ones_page, workspace, nsplit_out, stream);
}

static int wgrad_pw_reduce_launch(const float* workspace, int nsplit, int Cin, float* dW, int Cin_total, int co_n, int ci0,
                                  int accumulate, float* dbias, int accumulate_bias, void* stream) {
  RTP_CHECK_ARG(workspace && dW && nsplit >= 1 && co_n >= 1 && co_n <= 128 && ci0 >= 0 && ci0 + Cin <= Cin_total,
                "rtp_wgrad_pw_reduce: bad args");
  const int NX = nx_of(Cin, dbias != nullptr);
  const int total = co_n * Cin + (dbias ? co_n : 0);
  wgrad_pw_reduce_kernel<<<ceil_div(total, 32), 256, 0, (cudaStream_t)stream>>>(workspace, nsplit, NX, dW, Cin_total, co_n, ci0, Cin,
                                                                                 accumulate, dbias, dbias ? (Cin + 7) / 8 * 8 : -1,
                                                                                 accumulate_bias);
  RTP_LAUNCH_CHECK();
}
extern "C" int rtp_wgrad_pw_reduce(const float* workspace, int32_t nsplit, int32_t Cin, float* dW, int32_t Cin_total, int32_t co_n,
                                   int32_t ci0, int32_t accumulate, void* stream) {
  return wgrad_pw_reduce_launch(workspace, nsplit, Cin, dW, Cin_total, co_n, ci0, accumulate, nullptr, 0, stream);
}
extern "C" int rtp_wgrad_pw_bias_reduce(const float* workspace, int32_t nsplit, int32_t Cin, float* dW, int32_t Cin_total, int32_t co_n,
                                        int32_t ci0, int32_t accumulate, float* dbias, int32_t accumulate_bias, void* stream) {
  RTP_CHECK_ARG(dbias, "rtp_wgrad_pw_bias_reduce: null dbias");
  return wgrad_pw_reduce_launch(workspace, nsplit, Cin, dW, Cin_total, co_n, ci0, accumulate, dbias, accumulate_bias, stream);
}
