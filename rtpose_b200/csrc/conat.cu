// conat.cu — the final `cat(x0, up(x1), up(x2), up(x3)) -> Conv3d 1x1x1` of HRNet3D (backbones/hrnet3d.py:37-42, trilinear
// align_corners=True upsampling of the three lower-resolution branches to the full grid, concat, final_conv) as ONE
// streaming tcgen05 GEMM whose A operand is assembled in shared memory: the concat never exists in HBM.
//
// GEMM per 128-position tile of the padded full-resolution volume: D[128 x Cout] = A[128 x K] * W^T, K = C0 + sum C_j.
//   * chunks of x0 (K-major SWIZZLE_NONE canonical layout = the P8 layout itself) arrive by bulk async copies (warp 0);
//   * chunks of up(x_j) are INTERPOLATED into the stage by 8 warps: trilinear interpolation is separable, so for every low
//     term the warps first blend the 4 (z, x) corner rows of the <= 4 output rows a tile touches into shared memory (fp32;
//     4 loads per blended vector), then every (position, chunk) is a y-interpolation of two shared-memory vectors, rounded
//     to bf16 and stored as one 16-byte row of the operand (generic-proxy stores + fence.proxy.async + mbarrier arrive);
//   * warp 1 issues K/16 UMMAs (M128 x N=Cout x K16) into one of two TMEM accumulators; warps 2-5 drain (bias, bf16,
//     coalesced 16-byte stores) while the next tile is being assembled.
// Before: a 1x1 conv per branch (the full-resolution one wrote a Cout-channel tensor that was read again), then a
// fuse_sum pass that interpolated Cout = 128 channels per low term (384 channel-interpolations per voxel, issue-bound:
// 0.69 ms + 0.25 ms per 16 frames).  Here 160 channel-interpolations per voxel, x0 read once, the result written once.
// Interpolation arithmetic: fuse.cu's (ATen upsample_trilinear3d, align_corners=True).
//
// Roofline: HBM — x0 (C0/8 chunks) read + Cout/8 chunks written per voxel (64 B + 256 B at 32 -> 128 channels); the low
// terms are L2-resident.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 14 * 32;  // warp 0: bulk producer (x0), warp 1: MMA, warps 2-5: epilogue, warps 6-13: interpolators
constexpr int kInterpWarp0 = 6;  // two interpolator groups of 128 threads (warps 6-9, 10-13)
constexpr int kStages = 2;
constexpr int kMaxRows = 4;  // output rows (z, x) a 128-position tile may touch (Yp >= 43)

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

struct CT {
  P8 x0, out, low[3];
  int n_low, c8_x0, c8_low[3];
  const bf16* w;
  const float* bias;
  int K, NP, out_c8, relu;
  int npos, ntile, nunits;
  uint32_t stage_bytes, w_bytes;
  uint32_t sc_off[3];  // float4 offset of each term's blended rows in a group's scratch: [chunk][max_rows][Yl][2]
  uint32_t sc_group;   // float4s of one interpolator group's scratch
  uint32_t max_rows;   // output rows a 128-position tile may touch: 3 when Yp >= 64, else 4
};

__device__ __forceinline__ void bar_sync_group(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }  // ids 1..4

template <int MAXR>  // output rows a tile may touch (3 or 4)
__global__ void __launch_bounds__(kThreads, 1) conat_kernel(const __grid_constant__ CT p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_w, bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* wsm = smem;
  uint8_t* stages = smem + p.w_bytes;
  float4* scratch_all = reinterpret_cast<float4*>(smem + p.w_bytes + kStages * p.stage_bytes);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1 + 4); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_w, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 128); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ---------------------------------------------------------------- x0 chunks: one 2 KB bulk copy per chunk and tile
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, p.w_bytes);
      bulk_g2s(wsm, p.w, p.w_bytes, &bar_w);
    }
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int tile = u % p.ntile, n = u / p.ntile;
      const int s = it % kStages;
      if (lane == 0) {
        mbar_wait(&bar_empty[s], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], (uint32_t)p.c8_x0 * 2048u);
      }
      __syncwarp();
      const bf16* src = p.x0.ptr + (int64_t)n * p.x0.n_stride + (int64_t)tile * 128 * 8;
      uint8_t* dst = stages + (size_t)s * p.stage_bytes;
      for (int c = lane; c < p.c8_x0; c += 32) bulk_g2s(dst + (size_t)c * 2048, src + (int64_t)c * p.x0.c_stride, 2048, &bar_full[s]);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue
    const uint32_t idesc = idesc_bf16(128, p.NP, 0, 0);
    const uint32_t a_lo_c = (2048u >> 4) << 16, b_lo_c = (uint32_t)p.NP << 16, hi = (128u >> 4) | (1u << 14);
    const uint32_t stage0 = smem_u32(stages), w0 = smem_u32(wsm);
    mbar_wait(&bar_w, 0);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int s = it % kStages, buf = it & 1;
      mbar_wait(&bar_acc_empty[buf], ((it >> 1) & 1) ^ 1);
      mbar_wait(&bar_full[s], (it / kStages) & 1);
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
  } else if (warp < kInterpWarp0) {
    // ---------------------------------------------------------------- epilogue (one group of four warps, both accumulators)
    const int lane_q = warp & 3;
    const int r = lane_q * 32 + lane;
    const int Yp = p.out.Yp, Xp = p.out.Xp;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16) + buf * p.NP;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int q = tile * 128 + r;
      const int yp = q % Yp, xp = (q / Yp) % Xp;
      const bool ok = q < p.npos && xp >= 1 && xp <= p.out.X && yp >= 1 && yp <= p.out.Y;
      bf16* out_row = p.out.ptr + (int64_t)n * p.out.n_stride + (int64_t)q * 8;
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      fence_after_sync();
      for (int c32 = 0; c32 * 32 < p.NP; ++c32) {
        uint32_t v[32];
        const bool wide = c32 * 32 + 32 <= p.NP;
        if (wide) tmem_ld32(trow + c32 * 32, v); else tmem_ld16(trow + c32 * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
        tmem_ld_wait();
        if (c32 * 32 + 32 >= p.NP) {
          fence_before_sync();
          mbar_arrive(&bar_acc_empty[buf]);
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
          stg16(out_row + ch * p.out.c_stride, pack8(f));
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- interpolators: chunks of up(x_j)
    // Two groups of four warps; group g assembles the tiles with it & 1 == g into stage g (kStages == 2), with its own
    // blend rows and named barriers, so the global-load latency of one group's blend phase overlaps the other's arithmetic.
    const int g = (warp - kInterpWarp0) >> 2;
    const int t = tid - (kInterpWarp0 + 4 * g) * 32;  // 0..127
    const int Yp = p.out.Yp, Xp = p.out.Xp, Z = p.out.Z;
    float4* scratch = scratch_all + (size_t)g * p.sc_group;
    uint8_t* stage = stages + (size_t)g * p.stage_bytes;
    constexpr int MR = MAXR;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      if ((int)(it & 1) != g) continue;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int q0 = tile * 128;
      const int row0 = q0 / Yp;
      const int nrows = (q0 + 127) / Yp - row0 + 1;
      // ---- phase 1: blend the (z, x) corner rows of every output row of the tile (scratch is free: barrier B of the last tile)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j >= p.n_low) break;
        const P8& l = p.low[j];
        const int Yl = l.Y;
        const int yl = t % Yl, slot = t / Yl, nslots = 128 / Yl;  // lanes along yl: coalesced row reads
        if (slot >= nslots) continue;
        const float szj = ac_scale(l.Z, Z), sxj = ac_scale(l.X, p.out.X);
        // per output row of the tile: corner-row offsets and weights (rows outside the volume / in the pad ring: skipped)
        int o00[MAXR], o01[MAXR], o10[MAXR], o11[MAXR];  // element offsets inside a (sample, chunk) volume: < 2^31
        float wz0[MAXR], wz1[MAXR], wx0[MAXR], wx1[MAXR];
        bool rv[MAXR];
#pragma unroll
        for (int rho = 0; rho < MAXR; ++rho) {
          const int R = row0 + rho, z = R / Xp, xp = R - z * Xp;
          rv[rho] = rho < nrows && z < Z && xp >= 1 && xp <= p.out.X;
          const Axis az = ac_axis(rv[rho] ? z : 0, l.Z, szj), ax = ac_axis(rv[rho] ? xp - 1 : 0, l.X, sxj);
          o00[rho] = (int)l.voxel(az.i0, ax.i0, yl); o01[rho] = (int)l.voxel(az.i0, ax.i1, yl);
          o10[rho] = (int)l.voxel(az.i1, ax.i0, yl); o11[rho] = (int)l.voxel(az.i1, ax.i1, yl);
          wz0[rho] = az.w0; wz1[rho] = az.w1; wx0[rho] = ax.w0; wx1[rho] = ax.w1;
        }
        const bf16* lb = l.ptr + (int64_t)n * l.n_stride;
        float4* sc = scratch + p.sc_off[j];
        const int c8 = p.c8_low[j];
        for (int c = slot; c < c8; c += nslots) {  // a thread owns chunk c of all rows: 4 loads per row in flight together
          const bf16* b = lb + (int64_t)c * l.c_stride;
          uint4 v[MAXR][4];
#pragma unroll
          for (int rho = 0; rho < MAXR; ++rho) {
            if (!rv[rho]) continue;
            v[rho][0] = ldg16(b + o00[rho]); v[rho][1] = ldg16(b + o01[rho]);
            v[rho][2] = ldg16(b + o10[rho]); v[rho][3] = ldg16(b + o11[rho]);
          }
#pragma unroll
          for (int rho = 0; rho < MAXR; ++rho) {
            if (!rv[rho]) continue;
            float f00[8], f01[8], f10[8], f11[8], o[8];
            unpack8(v[rho][0], f00); unpack8(v[rho][1], f01); unpack8(v[rho][2], f10); unpack8(v[rho][3], f11);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = wz0[rho] * (wx0[rho] * f00[e] + wx1[rho] * f01[e]) + wz1[rho] * (wx0[rho] * f10[e] + wx1[rho] * f11[e]);
            float4* d = sc + ((c * MR + rho) * Yl + yl) * 2;
            d[0] = make_float4(o[0], o[1], o[2], o[3]);
            d[1] = make_float4(o[4], o[5], o[6], o[7]);
          }
        }
      }
      bar_sync_group(1 + 2 * g);
      // ---- phase 2: y interpolation of every chunk of this thread's position into the operand stage
      mbar_wait(&bar_empty[g], ((it >> 1) & 1) ^ 1);
      {
        const int q = q0 + t;
        const int R = q / Yp, yp = q - R * Yp, rho = R - row0;
        const int z = R / Xp, xp = R - z * Xp;
        const bool valid = q < p.npos && xp >= 1 && xp <= p.out.X && yp >= 1 && yp <= p.out.Y;
        uint8_t* arow = stage + (size_t)p.c8_x0 * 2048 + (size_t)t * 16;
        int cb = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j >= p.n_low) break;
          const int Yl = p.low[j].Y, c8 = p.c8_low[j];
          const Axis ay = ac_axis(valid ? yp - 1 : 0, Yl, ac_scale(Yl, p.out.Y));
          const float4* sc = scratch + p.sc_off[j] + (size_t)(valid ? rho : 0) * Yl * 2;
          const int i0 = ay.i0 * 2, i1 = ay.i1 * 2, cstep = MR * Yl * 2;
#pragma unroll 2
          for (int c = 0; c < c8; ++c) {
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
              const float4* rowp = sc + c * cstep;
              const float4 a0 = rowp[i0], a1 = rowp[i0 + 1], b0 = rowp[i1], b1 = rowp[i1 + 1];
              float o[8];
              o[0] = ay.w0 * a0.x + ay.w1 * b0.x; o[1] = ay.w0 * a0.y + ay.w1 * b0.y;
              o[2] = ay.w0 * a0.z + ay.w1 * b0.z; o[3] = ay.w0 * a0.w + ay.w1 * b0.w;
              o[4] = ay.w0 * a1.x + ay.w1 * b1.x; o[5] = ay.w0 * a1.y + ay.w1 * b1.y;
              o[6] = ay.w0 * a1.z + ay.w1 * b1.z; o[7] = ay.w0 * a1.w + ay.w1 * b1.w;
              pk = pack8(o);
            }
            *reinterpret_cast<uint4*>(arow + (size_t)(cb + c) * 2048) = pk;
          }
          cb += c8;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[g]);
      bar_sync_group(2 + 2 * g);  // every thread of the group is done reading the blend rows
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

bool plan(const rtp_conat_desc* d, CT& k, size_t& smem) {
  const int K = d->K, NP = d->NP;
  if (K % 16 != 0 || K < 16 || NP % 16 != 0 || NP < 16 || NP > 128) return false;  // two accumulators in 256 TMEM columns
  if (d->n_low < 1 || d->n_low > 3) return false;
  if (d->x0.Y + 2 < 43) return false;  // a 128-position tile must touch <= kMaxRows rows
  int ksum = d->c_x0;
  if (d->c_x0 % 16 != 0 || d->c_x0 < 16) return false;
  size_t sc = 0;
  k.max_rows = d->x0.Y + 2 >= 64 ? 3 : kMaxRows;
  for (int j = 0; j < d->n_low; ++j) {
    if (d->c_low[j] % 16 != 0 || d->c_low[j] < 16) return false;
    if (d->low[j].Y < 1 || d->low[j].Y > 128) return false;  // a blend pass puts the lanes of a group along the low-res row
    k.sc_off[j] = (uint32_t)sc;
    sc += (size_t)(d->c_low[j] / 8) * k.max_rows * d->low[j].Y * 2;
    ksum += d->c_low[j];
  }
  if (ksum != K) return false;
  k.sc_group = (uint32_t)sc;
  k.w_bytes = (uint32_t)K * NP * 2;
  k.stage_bytes = (uint32_t)K * 256;
  smem = (size_t)k.w_bytes + (size_t)kStages * k.stage_bytes + 2 * sc * 16;
  return smem <= 225 * 1024;
}

}  // namespace

extern "C" int rtp_conat_supported(const rtp_conat_desc* d) {
  if (!d) return 0;
  CT k;
  size_t smem;
  return plan(d, k, smem) ? 1 : 0;
}

extern "C" int rtp_conat_fwd(const rtp_conat_desc* d, void* stream) {
  RTP_CHECK_ARG(d && d->x0.ptr && d->out.ptr && d->w, "rtp_conat_fwd: null argument");
  CT k;
  size_t smem;
  RTP_CHECK_ARG(plan(d, k, smem), "rtp_conat_fwd: unsupported shape (K=%d NP=%d n_low=%d Y=%d)", d->K, d->NP, d->n_low, d->x0.Y);
  RTP_CHECK_ARG(d->x0.N == d->out.N && d->x0.Z == d->out.Z && d->x0.X == d->out.X && d->x0.Y == d->out.Y,
                "rtp_conat_fwd: x0 / out geometry mismatch");
  const int64_t plane = (int64_t)(d->x0.X + 2) * (d->x0.Y + 2) * 8;
  RTP_CHECK_ARG(d->x0.c_stride == d->x0.Z * plane && d->out.c_stride == d->out.Z * plane,
                "rtp_conat_fwd: planes must be contiguous per chunk");
  RTP_CHECK_ARG(d->x0.C8 * 8 >= d->c_x0 && d->out_c8 >= 1 && d->out_c8 * 8 <= d->NP && d->out_c8 <= d->out.C8,
                "rtp_conat_fwd: bad channel counts");
  k.x0 = P8(d->x0); k.out = P8(d->out);
  k.n_low = d->n_low; k.c8_x0 = d->c_x0 / 8;
  for (int j = 0; j < 3; ++j) {
    k.c8_low[j] = 0;
    if (j >= d->n_low) { k.low[j] = P8(d->x0); continue; }
    RTP_CHECK_ARG(d->low[j].ptr && d->low[j].N == d->x0.N && d->low[j].C8 * 8 >= d->c_low[j] && d->low[j].Y >= 1 &&
                      d->low[j].X >= 1 && d->low[j].Z >= 1,
                  "rtp_conat_fwd: low[%d] mismatch", j);
    k.low[j] = P8(d->low[j]);
    k.c8_low[j] = d->c_low[j] / 8;
  }
  k.w = (const bf16*)d->w; k.bias = d->bias; k.K = d->K; k.NP = d->NP; k.out_c8 = d->out_c8; k.relu = d->relu;
  k.npos = d->x0.Z * (d->x0.X + 2) * (d->x0.Y + 2);
  k.ntile = (k.npos + 127) / 128;
  k.nunits = d->x0.N * k.ntile;
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  auto kern = k.max_rows == 3 ? conat_kernel<3> : conat_kernel<4>;
  if (smem > configured) {
    for (auto kf : {conat_kernel<3>, conat_kernel<4>}) {
      cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { rtp_set_error("rtp_conat_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    configured = smem;
  }
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}
