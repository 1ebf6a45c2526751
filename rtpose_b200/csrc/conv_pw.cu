// conv_pw.cu — pointwise (1x1x1) conv3d, forward and dgrad, on tcgen05: a streaming GEMM over the linear positions of
// a P8 tensor.  In P8 the A operand of a 128-position tile is already in the SWIZZLE_NONE K-major canonical layout
// ([channel chunk][128 positions][8 channels] = one contiguous 2 KB run per chunk), so the producer is just K/8 bulk
// async copies per tile; weights stay resident in shared memory; two TMEM accumulators ping-pong between the MMA warp
// and the epilogue warps.  Used for the HRNet fuse-layer 1x1 convs (hr_util/hr3d.py:147-157), the per-branch slices of
// the final 192->128 conv (backbones/hrnet3d.py:20,41) and their dgrads.
//
// Roofline: HBM (K/8 input chunks read + NP/8 output chunks written per position; e.g. 32 -> 128 channels moves
// 64 B + 256 B per voxel for 8 kFLOP).
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 320;  // warp 0 producer, warp 1 MMA, warps 2-5 / 6-9 epilogue groups (one TMEM accumulator each)
constexpr int kMaxStages = 8;

struct PW {
  P8 in, out, mask;
  const bf16* w;
  const float* bias;
  int K, NP, out_c8, relu, accumulate, has_mask;
  int npos;      // positions per sample: Z * Xp * Yp
  int ntile;     // 128-position tiles per sample
  int nunits;    // N * ntile
  int nstages;
  uint32_t stage_bytes, w_bytes;
};

__global__ void __launch_bounds__(kThreads, 1) conv_pw_kernel(const __grid_constant__ PW p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_w, bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.nstages, kch = p.K >> 3;
  uint8_t* wsm = smem;
  uint8_t* stages = smem + p.w_bytes;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_w, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 128); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // lane 0 runs the barrier protocol; the K/8 chunk copies of a tile are issued by K/8 lanes at once (a single
    // thread issuing 2 KB bulk copies back to back tops out near 8 GB/s per SM, see profiles/r01_bulk_probe.txt)
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, p.w_bytes);
      bulk_g2s(wsm, p.w, p.w_bytes, &bar_w);
    }
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int tile = u % p.ntile, n = u / p.ntile;
      const int s = it % S;
      if (lane == 0) {
        mbar_wait(&bar_empty[s], ((it / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], p.stage_bytes);
      }
      __syncwarp();
      const bf16* src = p.in.ptr + (int64_t)n * p.in.n_stride + (int64_t)tile * 128 * 8;
      uint8_t* dst = stages + (size_t)s * p.stage_bytes;
      for (int c = lane; c < kch; c += 32) bulk_g2s(dst + (size_t)c * 2048, src + (int64_t)c * p.in.c_stride, 2048, &bar_full[s]);
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128, p.NP, 0, 0);
    const uint32_t a_lo_c = (2048u >> 4) << 16, b_lo_c = (uint32_t)p.NP << 16, hi = (128u >> 4) | (1u << 14);
    const uint32_t stage0 = smem_u32(stages), w0 = smem_u32(wsm);
    mbar_wait(&bar_w, 0);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int s = it % S, buf = it & 1;
      mbar_wait(&bar_acc_empty[buf], ((it >> 1) & 1) ^ 1);
      mbar_wait(&bar_full[s], (it / S) & 1);
      fence_after_sync();
      const uint32_t a_lo = a_lo_c + ((stage0 + (uint32_t)s * p.stage_bytes) >> 4);
      const uint32_t b_lo = b_lo_c + (w0 >> 4);
      if (elect_one()) {
        for (int k16 = 0; k16 < (p.K >> 4); ++k16)
          mma_ss(tmem + buf * p.NP, ((uint64_t)hi << 32) | (a_lo + k16 * 256), ((uint64_t)hi << 32) | (b_lo + k16 * 2 * p.NP), idesc,
                 k16 ? 1u : 0u);
        mma_commit(&bar_empty[s]);
        mma_commit(&bar_acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    const int lane_q = warp & 3, grp = (warp - 2) >> 2;   // group g drains accumulator g (tiles with it&1 == g)
    const int r = lane_q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16) + grp * p.NP;
    const int Yp = p.out.Yp, Xp = p.out.Xp;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      if ((int)(it & 1) != grp) continue;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int q = tile * 128 + r;                        // linear position inside the sample (all planes)
      const int yp = q % Yp, xp = (q / Yp) % Xp;
      const bool ok = q < p.npos && xp >= 1 && xp <= p.out.X && yp >= 1 && yp <= p.out.Y;
      bf16* out_row = p.out.ptr + (int64_t)n * p.out.n_stride + (int64_t)q * 8;
      const bf16* mask_row = p.mask.ptr + (int64_t)n * p.mask.n_stride + (int64_t)q * 8;
      // mask / accumulate operands of the first 4 chunks are fetched before the accumulator wait
      uint4 pm[4], pa[4];
      const bool pre = ok && p.out_c8 <= 4;
      if (pre) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < p.out_c8 && p.has_mask) pm[c] = ldg16(mask_row + c * p.mask.c_stride);
          if (c < p.out_c8 && p.accumulate) pa[c] = *reinterpret_cast<const uint4*>(out_row + c * p.out.c_stride);
        }
      }
      mbar_wait(&bar_acc_full[grp], (it >> 1) & 1);
      fence_after_sync();
      for (int c32 = 0; c32 * 32 < p.NP; ++c32) {
        uint32_t v[32];
        const bool wide = c32 * 32 + 32 <= p.NP;
        if (wide) tmem_ld32(trow + c32 * 32, v); else tmem_ld16(trow + c32 * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
        tmem_ld_wait();
        if (c32 * 32 + 32 >= p.NP) {
          fence_before_sync();
          mbar_arrive(&bar_acc_empty[grp]);
        }
        if (!ok) continue;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int ch = c32 * 4 + h;
          if (ch >= p.out_c8 || (!wide && h >= 2)) continue;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[h * 8 + i]);
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch * 8 + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (p.has_mask) {
            float g[8];
            unpack8((pre && c32 == 0) ? pm[h] : ldg16(mask_row + ch * p.mask.c_stride), g);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = g[i] > 0.f ? f[i] : 0.f;
          }
          bf16* dst = out_row + ch * p.out.c_stride;
          if (p.accumulate) {
            float g[8];
            unpack8((pre && c32 == 0) ? pa[h] : *reinterpret_cast<const uint4*>(dst), g);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += g[i];
          }
          stg16(dst, pack8(f));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

bool plan(int K, int NP, int& stages, size_t& smem) {
  if (K % 16 != 0 || K < 16 || K > 256 || NP % 16 != 0 || NP < 16 || NP > 256) return false;
  const size_t w = (size_t)K * NP * 2, st = (size_t)K * 256;
  if (w + 2 * st > 220 * 1024) return false;
  stages = (int)((220 * 1024 - w) / st);
  if (stages > kMaxStages) stages = kMaxStages;
  smem = w + stages * st;
  return true;
}

}  // namespace

extern "C" int rtp_conv_pw_supported(int32_t K, int32_t NP) {
  int s;
  size_t m;
  return plan(K, NP, s, m) ? 1 : 0;
}

extern "C" int rtp_conv_pw(rtp_p8 in, rtp_p8 out, rtp_p8 mask, const void* w, const float* bias, int32_t K, int32_t NP,
                           int32_t out_c8, int32_t relu, int32_t accumulate, void* stream) {
  RTP_CHECK_ARG(in.ptr && out.ptr && w, "rtp_conv_pw: null argument");
  RTP_CHECK_ARG(in.N == out.N && in.Z == out.Z && in.X == out.X && in.Y == out.Y, "rtp_conv_pw: geometry mismatch");
  RTP_CHECK_ARG(in.C8 * 8 >= K, "rtp_conv_pw: input has %d channels, K=%d", in.C8 * 8, K);
  RTP_CHECK_ARG(out_c8 >= 1 && out_c8 * 8 <= NP + 7 && out_c8 <= out.C8, "rtp_conv_pw: bad out_c8");
  const int64_t plane = (int64_t)(in.X + 2) * (in.Y + 2) * 8;
  RTP_CHECK_ARG(in.c_stride == in.Z * plane && out.c_stride == out.Z * plane, "rtp_conv_pw: planes must be contiguous per chunk");
  PW k;
  size_t smem;
  RTP_CHECK_ARG(plan(K, NP, k.nstages, smem), "rtp_conv_pw: unsupported K=%d NP=%d", K, NP);
  k.in = P8(in); k.out = P8(out); k.mask = P8(mask); k.has_mask = mask.ptr != nullptr;
  if (k.has_mask) RTP_CHECK_ARG(mask.c_stride == out.c_stride, "rtp_conv_pw: mask layout mismatch");
  k.w = (const bf16*)w; k.bias = bias; k.K = K; k.NP = NP; k.out_c8 = out_c8; k.relu = relu; k.accumulate = accumulate;
  k.npos = in.Z * (in.X + 2) * (in.Y + 2);
  k.ntile = (k.npos + 127) / 128;
  k.nunits = in.N * k.ntile;
  k.stage_bytes = (uint32_t)K * 256; k.w_bytes = (uint32_t)K * NP * 2;
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_conv_pw: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  conv_pw_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}
