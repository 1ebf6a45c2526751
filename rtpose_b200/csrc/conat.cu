// conat.cu — the final `cat(x0, up(x1), up(x2), up(x3)) -> Conv3d 1x1x1` of HRNet3D (backbones/hrnet3d.py:37-42, trilinear
// align_corners=True upsampling of the three lower-resolution branches to the full grid, concat, final_conv) as ONE
// streaming tcgen05 GEMM whose A operand is assembled in shared memory: the concat never exists in HBM.
//
// GEMM per tile of RPT = 128 / Y whole output rows (RPT * Y <= 128 positions, pad ring excluded):
// D[128 x Cout] = A[128 x K] * W^T, K = C0 + sum C_j.
//   * chunks of x0 (K-major SWIZZLE_NONE canonical layout = the P8 layout itself) arrive by one bulk async copy per
//     (chunk, row) (warp 0);
//   * chunks of up(x_j) are INTERPOLATED into the stage by two groups of four warps (group g builds the tiles with
//     it & 1 == g into stage g): trilinear interpolation is separable, so a group first blends, for every low term, the 4
//     (z, x) corner rows of each output row of the tile into shared memory (fp32; 4 global loads per blended vector), then
//     every thread owns one position of the tile — its y-interpolation indices and weights are kernel constants held in
//     registers — and writes, per chunk, the blend of two shared-memory vectors as one 16-byte bf16 row of the operand
//     (generic-proxy stores + fence.proxy.async + mbarrier arrive);
//   * warp 1 issues K/16 UMMAs (M128 x N=Cout x K16) into one of two TMEM accumulators; warps 2-5 drain (bias, bf16,
//     coalesced 16-byte stores) while the next tiles are being assembled.
// Before: a 1x1 conv per branch (the full-resolution one wrote a Cout-channel tensor that was read again), then a
// fuse_sum pass that interpolated Cout = 128 channels per low term (384 channel-interpolations per voxel, issue-bound:
// 0.68 ms + 0.25 ms per 16 frames).  Here 160 channel-interpolations per voxel, x0 read once, the result written once.
// Interpolation arithmetic: fuse.cu's (ATen upsample_trilinear3d, align_corners=True).
//
// Roofline: HBM — x0 (C0/8 chunks) read + Cout/8 chunks written per voxel (64 B + 256 B at 32 -> 128 channels); the low
// terms are L2-resident.  Measured: instruction-issue bound in the interpolator warps (profiles/r02_ncu_conat.txt).
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 14 * 32;  // warp 0: bulk producer (x0), warp 1: MMA, warps 2-5: epilogue, warps 6-13: interpolators
constexpr int kInterpWarp0 = 6;    // two interpolator groups of 128 threads (warps 6-9, 10-13)
constexpr int kStages = 2;
// output rows per tile: RPT = 128 / Y <= 4 (Y >= 32)

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
  int RPT;             // output rows per tile
  int nrow;            // Z * X output rows per sample
  int ntile, nunits;
  uint32_t stage_bytes, w_bytes;
  uint32_t sc_off[3];  // float4 offset of each term's blended rows in a group's scratch: [chunk][RPT][2 halves][Yl]
  uint32_t sc_group;   // float4s of one interpolator group's scratch
};

__device__ __forceinline__ void bar_sync_group(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }  // ids 1..4
// waits of the warps that are NOT on the critical path (producer, epilogue): poll with a pause, the issue slots belong to
// the interpolators (ncu of the first version: 23 % of all executed instructions were barrier polls)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

__global__ void __launch_bounds__(kThreads, 1) conat_kernel(const __grid_constant__ CT p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_w, bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* wsm = smem;
  uint8_t* stages = smem + p.w_bytes;
  float4* scratch_all = reinterpret_cast<float4*>(smem + p.w_bytes + kStages * p.stage_bytes);
  const int Y = p.out.Y, X = p.out.X, RPT = p.RPT;

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
    // ---------------------------------------------------------------- x0 chunks: one bulk copy per (chunk, row) and tile
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_w, p.w_bytes);
      bulk_g2s(wsm, p.w, p.w_bytes, &bar_w);
    }
    const uint32_t row_bytes = (uint32_t)Y * 16u;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int tile = u % p.ntile, n = u / p.ntile;
      const int s = it % kStages;
      const int R0 = tile * RPT, nvalid = min(RPT, p.nrow - R0);
      if (lane == 0) {
        mbar_wait_relaxed(&bar_empty[s], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], (uint32_t)(p.c8_x0 * nvalid) * row_bytes);
      }
      __syncwarp();
      const bf16* src = p.x0.ptr + (int64_t)n * p.x0.n_stride;
      uint8_t* dst = stages + (size_t)s * p.stage_bytes;
      for (int i = lane; i < p.c8_x0 * nvalid; i += 32) {
        const int c = i / nvalid, rho = i - c * nvalid;
        const int R = R0 + rho, z = R / X, x = R - z * X;
        bulk_g2s(dst + (size_t)c * 2048 + (size_t)rho * row_bytes, src + (int64_t)c * p.x0.c_stride + p.x0.voxel(z, x, 0), row_bytes,
                 &bar_full[s]);
      }
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
      mbar_wait_relaxed(&bar_full[s], (it / kStages) & 1);
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
    const int rho = r / Y, y = r - rho * Y;  // this thread's row of the tile and position in it: kernel constants
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16) + buf * p.NP;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int R = tile * RPT + rho, z = R / X, x = R - z * X;
      const bool ok = rho < RPT && R < p.nrow;
      bf16* out_row = p.out.ptr + (int64_t)n * p.out.n_stride + p.out.voxel(z, x, y);
      mbar_wait_relaxed(&bar_acc_full[buf], (it >> 1) & 1);
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
    const int g = (warp - kInterpWarp0) >> 2;
    const int t = tid - (kInterpWarp0 + 4 * g) * 32;  // 0..127
    const int Z = p.out.Z;
    float4* scratch = scratch_all + (size_t)g * p.sc_group;
    uint8_t* stage = stages + (size_t)g * p.stage_bytes;
    // ---- kernel constants of this thread
    // phase 2: position t of a tile = (row rho2, y2); per term the y-interpolation (two blended vectors, two weights)
    const int rho2 = t / Y, y2 = t - rho2 * Y;
    const bool pos_ok = rho2 < RPT;
    int p2_i0[3], p2_i1[3];   // float4 offsets inside a term's scratch for chunk 0
    float p2_w0[3], p2_w1[3];
    // phase 1: lane position along the low-resolution row and first (row, chunk) item of this thread, per term
    int p1_yl[3], p1_slot[3], p1_nslots[3];
    float p1_sz[3], p1_sx[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int Yl = j < p.n_low ? p.low[j].Y : 1;
      const Axis ay = ac_axis(pos_ok ? y2 : 0, Yl, ac_scale(Yl, Y));
      p2_i0[j] = (pos_ok ? rho2 : 0) * 2 * Yl + ay.i0;
      p2_i1[j] = (pos_ok ? rho2 : 0) * 2 * Yl + ay.i1;
      p2_w0[j] = ay.w0; p2_w1[j] = ay.w1;
      p1_yl[j] = t % Yl; p1_slot[j] = t / Yl; p1_nslots[j] = 128 / Yl;
      p1_sz[j] = j < p.n_low ? ac_scale(p.low[j].Z, Z) : 0.f;
      p1_sx[j] = j < p.n_low ? ac_scale(p.low[j].X, X) : 0.f;
    }
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      if ((int)(it & 1) != g) continue;
      const int tile = u % p.ntile, n = u / p.ntile;
      const int R0 = tile * RPT, z0 = R0 / X, x0r = R0 - z0 * X;
      // ---- phase 1: blend the (z, x) corner rows of every output row of the tile (scratch is free: barrier B of the last tile)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j >= p.n_low) break;
        if (p1_slot[j] >= p1_nslots[j]) continue;
        const P8& l = p.low[j];
        const int Yl = l.Y, yl = p1_yl[j];
        const bf16* lb = l.ptr + (int64_t)n * l.n_stride;
        float4* sc = scratch + p.sc_off[j];
        const int items = p.c8_low[j] * RPT;  // (chunk, row) pairs, row fastest
        for (int m0 = p1_slot[j]; m0 < items; m0 += 2 * p1_nslots[j]) {
          uint4 v[2][4];
          float wz0[2], wz1[2], wx0[2], wx1[2];
          int dst[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int m = m0 + k * p1_nslots[j];
            dst[k] = -1;
            if (m >= items) continue;
            const int c = m / RPT, rho = m - c * RPT;
            int x = x0r + rho, z = z0;
            if (x >= X) { x -= X; ++z; }
            if (z >= Z) continue;  // rows past the last one (the final tile of a sample)
            const Axis az = ac_axis(z, l.Z, p1_sz[j]), ax = ac_axis(x, l.X, p1_sx[j]);
            const bf16* b = lb + (int64_t)c * l.c_stride;
            v[k][0] = ldg16(b + l.voxel(az.i0, ax.i0, yl));
            v[k][1] = ldg16(b + l.voxel(az.i0, ax.i1, yl));
            v[k][2] = ldg16(b + l.voxel(az.i1, ax.i0, yl));
            v[k][3] = ldg16(b + l.voxel(az.i1, ax.i1, yl));
            wz0[k] = az.w0; wz1[k] = az.w1; wx0[k] = ax.w0; wx1[k] = ax.w1;
            dst[k] = (c * RPT + rho) * 2 * Yl + yl;  // [chunk][row][half][yl]: lanes along yl are 16 bytes apart (no bank conflicts)
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (dst[k] < 0) continue;
            float f00[8], f01[8], f10[8], f11[8], o[8];
            unpack8(v[k][0], f00); unpack8(v[k][1], f01); unpack8(v[k][2], f10); unpack8(v[k][3], f11);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = wz0[k] * (wx0[k] * f00[e] + wx1[k] * f01[e]) + wz1[k] * (wx0[k] * f10[e] + wx1[k] * f11[e]);
            sc[dst[k]] = make_float4(o[0], o[1], o[2], o[3]);
            sc[dst[k] + Yl] = make_float4(o[4], o[5], o[6], o[7]);
          }
        }
      }
      bar_sync_group(1 + 2 * g);
      // ---- phase 2: y interpolation of every chunk of this thread's position into the operand stage
      mbar_wait(&bar_empty[g], ((it >> 1) & 1) ^ 1);
      if (pos_ok && R0 + rho2 < p.nrow) {
        uint8_t* arow = stage + (size_t)p.c8_x0 * 2048 + (size_t)t * 16;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j >= p.n_low) break;
          const int c8 = p.c8_low[j], Yl = p.low[j].Y, cstep = RPT * Yl * 2;
          const float4* s0 = scratch + p.sc_off[j] + p2_i0[j];
          const float4* s1 = scratch + p.sc_off[j] + p2_i1[j];
          const float w0 = p2_w0[j], w1 = p2_w1[j];
#pragma unroll 2
          for (int c = 0; c < c8; ++c) {
            const float4 a0 = s0[c * cstep], a1 = s0[c * cstep + Yl], b0 = s1[c * cstep], b1 = s1[c * cstep + Yl];
            float o[8];
            o[0] = w0 * a0.x + w1 * b0.x; o[1] = w0 * a0.y + w1 * b0.y; o[2] = w0 * a0.z + w1 * b0.z; o[3] = w0 * a0.w + w1 * b0.w;
            o[4] = w0 * a1.x + w1 * b1.x; o[5] = w0 * a1.y + w1 * b1.y; o[6] = w0 * a1.z + w1 * b1.z; o[7] = w0 * a1.w + w1 * b1.w;
            *reinterpret_cast<uint4*>(arow + (size_t)c * 2048) = pack8(o);
          }
          arow += (size_t)c8 * 2048;
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
  if (d->x0.Y < 32 || d->x0.Y > 128) return false;  // a tile is RPT = 128 / Y <= 4 whole rows
  k.RPT = 128 / d->x0.Y;
  int ksum = d->c_x0;
  if (d->c_x0 % 16 != 0 || d->c_x0 < 16) return false;
  size_t sc = 0;
  for (int j = 0; j < d->n_low; ++j) {
    if (d->c_low[j] % 16 != 0 || d->c_low[j] < 16) return false;
    if (d->low[j].Y < 1 || d->low[j].Y > 128) return false;  // a blend pass puts the lanes of a group along the low-res row
    k.sc_off[j] = (uint32_t)sc;
    sc += (size_t)(d->c_low[j] / 8) * k.RPT * d->low[j].Y * 2;
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
  k.nrow = d->x0.Z * d->x0.X;
  k.ntile = (k.nrow + k.RPT - 1) / k.RPT;
  k.nunits = d->x0.N * k.ntile;
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_conat_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  conat_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}
