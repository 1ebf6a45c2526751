// wgrad_generic.cu — conv3d weight gradient on tcgen05 for any tap list / stride.
//
//   dW[tap][ci][co] = sum_rows X[row + tap][ci] * dY[row][co]
//
// GEMM view: K = rows (voxels), both operands MN-major (channels contiguous, K strided) — exactly how the P8
// layout stores them.  M = 128 stacks (tap, 8-channel input chunk) pairs ("m-blocks" of 16 pairs), N = NP output
// channels.  Each CTA owns a contiguous range of 64-row tiles (split-K) and a group of m-blocks whose fp32
// accumulators [128 x NP] all stay resident in TMEM for the whole range; the partial result goes to a
// workspace and rtp_wgrad_reduce sums the splits in a fixed order (deterministic) into the reference's
// [Cout][Cin][taps] fp32 layout.
#include <climits>
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kStages = 4;
constexpr int kProducers = 256;           // 8 gather warps: thread = (row of the tile, quarter of the pair slots)
constexpr int kThreads = kProducers + 32;  // + the MMA warp
constexpr int kTileK = 64;  // rows per pipeline item

struct WgradK {
  P8 x, dy;
  int Cin, NP, ntaps;
  int8_t tz[RTP_MAX_TAPS], tx[RTP_MAX_TAPS], ty[RTP_MAX_TAPS];
  int16_t tc[RTP_MAX_TAPS];  // first 8-channel chunk of X read by the tap (space-to-depth views: parity group * Cin/8)
  int RZ, RX, RY, IS;
  int64_t total_rows;
  int ntiles, tiles_per_split;
  int npairs;        // ntaps * Cin/8
  int nblocks;       // ceil(npairs / 16)
  int blocks_per_cta;
  int tmem_cols, col_stride;
  float* partial;    // [nsplit][nblocks][128][NP]
};

__global__ void __launch_bounds__(kThreads) wgrad_generic_kernel(const __grid_constant__ WgradK p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_acc;
  __shared__ uint32_t tmem_base_s;
  // per (tap, chunk) pair: element offset from the row's voxel (or INT_MIN when the chunk does not exist) and the z tap.
  // Computed once: the gather loop below then costs one shared load + one add per 16-byte cp.async instead of ~30
  // integer instructions (the kernel runs at 2 CTAs / SM, so it was bound by exactly that dependent ALU chain).
  __shared__ int2 s_tab[RTP_MAX_TAPS * 32 + 16];  // {offset, z tap}; padded with INT_MIN up to a whole m-block

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_bytes = 16 * kTileK * 16;            // [16 pairs][64 rows][16 B]
  const uint32_t b_bytes = (p.NP / 8) * kTileK * 16;    // [NP/8][64 rows][16 B]
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int split = blockIdx.x;
  const int blk0 = blockIdx.y * p.blocks_per_cta;
  const int nblk = min(p.blocks_per_cta, p.nblocks - blk0);
  const int tile0 = split * p.tiles_per_split;
  const int tile1 = min(p.ntiles, tile0 + p.tiles_per_split);
  const int nitems = max(0, tile1 - tile0) * nblk;
  const int kch = p.Cin >> 3;

  for (int pair = tid; pair < p.nblocks * 16 && pair < RTP_MAX_TAPS * 32 + 16; pair += kThreads) {
    int2 e = make_int2(INT_MIN, 0);
    if (pair < p.npairs) {
      const int tap = pair / kch, c = pair - tap * kch;
      const int cc = p.tc[tap] + c;
      const int64_t off = (((int64_t)p.tz[tap] * p.x.Xp + p.tx[tap]) * p.x.Yp + p.ty[tap]) * 8 + (int64_t)cc * p.x.c_stride;
      e = make_int2(cc < p.x.C8 ? (int)off : INT_MIN, p.tz[tap]);
    }
    s_tab[pair] = e;
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], kProducers);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == kProducers / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < kProducers / 32) {
    const int r = tid;
    const int pos = r & (kTileK - 1);  // this thread always serves the same row of the tile
    const int half = r >> 6;           // and the pair slots half, half+4, half+8, half+12 ("half" = quarter 0..3)

    const uint32_t smem32 = smem_u32(smem) + (uint32_t)((half * kTileK + pos) * 16);  // this thread's slot in pair `half`
    const int lane = tid & 31;
    auto publish = [&](int item) {
      fence_proxy_async();
      mbar_arrive(&bar_full[item & (kStages - 1)]);
    };
    static_assert((kStages & (kStages - 1)) == 0, "kStages must be a power of two");
    // Nested (tile, m-block) loops: the row decode (32-bit, once per tile) and the item -> (tile, block) mapping cost no
    // divisions per item — with 2 CTAs x 4 producer warps per SM this loop's dependent integer chain IS the kernel's
    // critical path (ncu: 55 % of the stall samples inside it, MMA warp idle).
    int item = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      const int64_t L = (int64_t)tile * kTileK + pos;
      const bool row_ok = L < p.total_rows;
      int n = 0, rx = 0, ry = 0, rz = 0;
      if (row_ok) {
        uint32_t q = (uint32_t)L;  // total_rows < 2^31 (checked on the host)
        ry = (int)(q % (uint32_t)p.RY); q /= (uint32_t)p.RY;
        rx = (int)(q % (uint32_t)p.RX); q /= (uint32_t)p.RX;
        rz = (int)(q % (uint32_t)p.RZ);
        n = (int)(q / (uint32_t)p.RZ);
      }
      const bf16* x_row = p.x.ptr + (int64_t)n * p.x.n_stride + p.x.voxel(rz * p.IS, rx * p.IS, ry * p.IS);
      const bf16* dy_row = p.dy.ptr + (int64_t)n * p.dy.n_stride + p.dy.voxel(rz, rx, ry);
      const int izb = rz * p.IS;
      for (int b = 0; b < nblk; ++b, ++item) {
        const int st = item & (kStages - 1);
        if (item >= kStages) {  // one lane per warp polls; the others park on the warp barrier
          if (lane == 0) mbar_wait(&bar_empty[st], ((item / kStages) - 1) & 1);
          __syncwarp();
        }
        const uint32_t sA = smem32 + (uint32_t)st * stage_bytes, sB = sA + a_bytes;
        const int2* tab = s_tab + (blk0 + b) * 16 + half;
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // pair slots half, half+4, ...: 4 KB apart in the stage
          const int2 e = tab[4 * i];
          const bool ok = row_ok && e.x != INT_MIN && (unsigned)(izb + e.y) < (unsigned)p.x.Z;
          cp_async16_s32(sA + i * (4 * kTileK * 16), ok ? (const void*)(x_row + e.x) : (const void*)p.x.ptr, ok);
        }
        for (int c = half; c < p.NP / 8; c += 4) {
          const bool ok = row_ok && c < p.dy.C8;
          cp_async16_s32(sB + (c - half) * (kTileK * 16), ok ? (const void*)(dy_row + (int64_t)c * p.dy.c_stride) : (const void*)p.dy.ptr, ok);
        }
        cp_async_commit();
        if (item >= kStages - 1) {
          cp_async_wait<kStages - 1>();
          publish(item - (kStages - 1));
        }
      }
    }
    cp_async_wait<0>();
    for (int it2 = (nitems >= kStages - 1 ? nitems - (kStages - 1) : 0); it2 < nitems; ++it2) publish(it2);

    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 partials (warps 0-3)
    if (warp < 4) {
    if (nitems > 0) {
      mbar_wait(&bar_acc, 0);
      fence_after_sync();
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int b = 0; b < nblk; ++b) {
      float* dst = p.partial + (((size_t)split * p.nblocks + blk0 + b) * 128 + r) * p.NP;
      for (int c16 = 0; c16 * 16 < p.NP; ++c16) {
        uint32_t v[16];
        if (nitems > 0) {
          tmem_ld16(trow + b * p.col_stride + c16 * 16, v);
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
  } else if (tid == kProducers) {
    const uint32_t idesc = idesc_bf16(128, p.NP, 1, 1);
    for (int item = 0; item < nitems; ++item) {
      const int st = item % kStages;
      const int b = item % nblk;
      const bool first_tile = item < nblk;
      mbar_wait(&bar_full[st], (item / kStages) & 1);
      fence_after_sync();
      const uint32_t sA = smem_u32(smem + (size_t)st * stage_bytes);
      const uint32_t sB = sA + a_bytes;
#pragma unroll
      for (int k16 = 0; k16 < kTileK / 16; ++k16) {
        // MN-major SWIZZLE_NONE: LBO = stride between 8-row K groups, SBO = stride between 8-element MN chunks
        const uint64_t ad = smem_desc(sA + k16 * 256, 128, kTileK * 16);
        const uint64_t bd = smem_desc(sB + k16 * 256, 128, kTileK * 16);
        mma_ss(tmem + b * p.col_stride, ad, bd, idesc, (!first_tile || k16) ? 1u : 0u);
      }
      mma_commit(&bar_empty[st]);
    }
    if (nitems > 0) mma_commit(&bar_acc);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kProducers / 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
  }
}

// partial [nsplit][nblocks][128][NP]  ->  dW[co][ci0 + ci][tap] = sum_split partial[..][(tap,ci)][n0 + co]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int Cin8, int NP,
                                                           int ntaps, int nblocks, float* __restrict__ dW, int Cin_total,
                                                           int co_n, int n0, int ci0, int ci_n, int accumulate) {
  // 32 outputs x 8 split lanes per block: lane l sums splits l, l+8, ... (independent loads in flight), then the 8
  // lane sums are combined in a fixed order -> deterministic, and ~8x shorter dependent chains than one thread/output
  __shared__ float sh[8][33];
  const int kch = Cin8 >> 3;
  const int64_t total = (int64_t)ntaps * ci_n * co_n;
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const size_t sstride = (size_t)nblocks * 128 * NP;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + o;
    float acc = 0.f;
    int co = 0, ci = 0, tap = 0;
    if (i < total) {
      co = (int)(i % co_n);
      const int64_t r = i / co_n;
      ci = (int)(r % ci_n);
      tap = (int)(r / ci_n);
      const int pair = tap * kch + (ci >> 3);
      const int blk = pair >> 4, m = (pair & 15) * 8 + (ci & 7);
      const float* src = partial + ((size_t)blk * 128 + m) * NP + n0 + co;
      for (int s = sl; s < nsplit; s += 8) acc += src[s * sstride];
    }
    sh[sl][o] = acc;
    __syncthreads();
    if (sl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += sh[l][o];
      float* d = dW + ((int64_t)co * Cin_total + ci0 + ci) * ntaps + tap;
      *d = accumulate ? *d + t : t;
    }
    __syncthreads();
  }
}

int cols_pow2(int n) {
  int c = 32;
  while (c < n) c *= 2;
  return c;
}

}  // namespace

extern "C" int64_t rtp_wgrad_workspace_bytes(int32_t Cin, int32_t NP, int32_t ntaps, int32_t nsplit) {
  const int npairs = ntaps * (Cin / 8);
  const int nblocks = (npairs + 15) / 16;
  return (int64_t)nsplit * nblocks * 128 * NP * 4;
}

extern "C" int rtp_wgrad(const rtp_wgrad_desc* d, void* stream) {
  RTP_CHECK_ARG(d != nullptr && d->x.ptr && d->dy.ptr && d->workspace, "rtp_wgrad: null argument");
  RTP_CHECK_ARG(d->Cin >= 8 && d->Cin % 8 == 0 && d->Cin <= 256, "rtp_wgrad: Cin=%d must be a multiple of 8, <= 256", d->Cin);
  RTP_CHECK_ARG(d->NP >= 16 && d->NP % 16 == 0 && d->NP <= 256, "rtp_wgrad: NP=%d must be a multiple of 16 <= 256", d->NP);
  RTP_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= RTP_MAX_TAPS && d->nsplit >= 1, "rtp_wgrad: bad ntaps/nsplit");
  RTP_CHECK_ARG(d->IS == 1 || d->IS == 2, "rtp_wgrad: IS must be 1 or 2");
  RTP_CHECK_ARG(d->RZ <= d->dy.Z && d->RX <= d->dy.X && d->RY <= d->dy.Y && d->x.N == d->dy.N, "rtp_wgrad: row grid exceeds dY");
  for (int t = 0; t < d->ntaps; ++t) {
    RTP_CHECK_ARG(d->tx[t] >= -1 && (d->RX - 1) * d->IS + d->tx[t] <= d->x.X && d->ty[t] >= -1 &&
                      (d->RY - 1) * d->IS + d->ty[t] <= d->x.Y,
                  "rtp_wgrad: tap %d leaves the padded input plane", t);
  }
  WgradK k;
  k.x = P8(d->x); k.dy = P8(d->dy);
  k.Cin = d->Cin; k.NP = d->NP; k.ntaps = d->ntaps;
  for (int t = 0; t < RTP_MAX_TAPS; ++t) { k.tz[t] = d->tz[t]; k.tx[t] = d->tx[t]; k.ty[t] = d->ty[t]; k.tc[t] = d->tc[t]; }
  k.RZ = d->RZ; k.RX = d->RX; k.RY = d->RY; k.IS = d->IS;
  k.total_rows = (int64_t)d->x.N * d->RZ * d->RX * d->RY;
  RTP_CHECK_ARG(k.total_rows < (1ll << 31), "rtp_wgrad: too many rows");
  k.ntiles = (int)((k.total_rows + kTileK - 1) / kTileK);
  k.tiles_per_split = (k.ntiles + d->nsplit - 1) / d->nsplit;
  k.npairs = d->ntaps * (d->Cin / 8);
  k.nblocks = (k.npairs + 15) / 16;
  k.col_stride = d->NP;
  k.blocks_per_cta = 512 / d->NP;
  if (k.blocks_per_cta > k.nblocks) k.blocks_per_cta = k.nblocks;
  k.tmem_cols = cols_pow2(k.blocks_per_cta * d->NP);
  k.partial = d->workspace;
  const size_t smem = (size_t)kStages * (16 * kTileK * 16 + (d->NP / 8) * kTileK * 16);
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  dim3 grid(d->nsplit, (k.nblocks + k.blocks_per_cta - 1) / k.blocks_per_cta);
  wgrad_generic_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_wgrad_reduce(const float* workspace, int32_t nsplit, int32_t Cin8, int32_t NP, int32_t ntaps, float* dW,
                                int32_t Cin_total, int32_t co_n, int32_t n0, int32_t ci0, int32_t ci_n, int32_t accumulate,
                                void* stream) {
  RTP_CHECK_ARG(workspace && dW && nsplit >= 1, "rtp_wgrad_reduce: null argument");
  RTP_CHECK_ARG(co_n >= 1 && n0 >= 0 && n0 + co_n <= NP && ci_n >= 1 && ci_n <= Cin8 && ci0 >= 0 && ci0 + ci_n <= Cin_total,
                "rtp_wgrad_reduce: bad channel ranges");
  const int npairs = ntaps * (Cin8 / 8);
  const int nblocks = (npairs + 15) / 16;
  const int64_t total = (int64_t)ntaps * ci_n * co_n;
  const int blocks = ceil_div(total, 32) > 4096 ? 4096 : ceil_div(total, 32);
  wgrad_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(workspace, nsplit, Cin8, NP, ntaps, nblocks, dW, Cin_total,
                                                               co_n, n0, ci0, ci_n, accumulate);
  RTP_LAUNCH_CHECK();
}
