// conv_generic.cu — general implicit-GEMM conv3d on tcgen05 (any tap list, stride 1/2, forward or dgrad).
//
// One CTA = one tile of 128 GEMM rows (output voxels).  Warps 0-3 (128 threads, thread == row) gather the A
// operand of each tap with 16-byte cp.async (zero-fill outside the volume) into the SWIZZLE_NONE K-major
// canonical layout [K/8][128 rows][8], and the tap's packed weights as B; warp 4 issues tcgen05.mma (one
// elected thread) through a kStages-deep mbarrier ring; the accumulator [128 x NP] fp32 lives in TMEM.  The
// same 4 warps then run the epilogue (tcgen05.ld -> bias / residual / ReLU / mask / accumulate -> bf16 16-byte
// stores into the P8 tensor).  Several CTAs are resident per SM, so one tile's gathers overlap another's MMAs.
//
// Roofline: tensor-bound in principle, but the per-tap re-gather (27x input amplification through the LSU)
// bounds it near 20 % of peak; the dominant 3x3x3 stride-1 shape goes through conv_k3s1.cu instead.
#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 160;

constexpr int kMaxClasses = 8;

// One "class" = one tap list over one row grid.  A launch may carry several (rtp_conv_multi): the 8 output-parity classes
// of a stride-2 dgrad differ only in these fields, and at low resolution each alone is a latency-bound handful of CTAs.
struct ConvClass {
  int8_t tz[RTP_MAX_TAPS], tx[RTP_MAX_TAPS], ty[RTP_MAX_TAPS], wt[RTP_MAX_TAPS];
  int ntaps, RZ, RX, RY, oz0, ox0, oy0;
  int tile0;  // first CTA of this class
  int64_t total_rows;
};

struct ConvK {
  P8 in, out, res, mask;
  const bf16* w;
  const float* bias;
  int Cin, NP, out_c8;
  int KC, nk;  // K per pipeline item (<= Cin) and items per tap: wide-K convs stream a tap in several chunks
  int IS, OS, relu, accumulate;
  int tmem_cols;
  int has_res, has_mask;
  int ncls;
  ConvClass cls[kMaxClasses];
};

template <int kStages>
__global__ void __launch_bounds__(kThreads) conv_generic_kernel(const __grid_constant__ ConvK p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_acc;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int kch = p.KC >> 3;                  // 8-channel chunks of K per item
  const uint32_t a_bytes = p.KC * 256;        // [kch][128][16 B]
  const uint32_t b_bytes = p.KC * p.NP * 2;   // [kch][NP][16 B]
  int ci = 0;
  while (ci + 1 < p.ncls && (int)blockIdx.x >= p.cls[ci + 1].tile0) ++ci;
  const ConvClass& cl = p.cls[ci];
  const int nitems = cl.ntaps * p.nk;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 128);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    // ------------------------------------------------------------------ producer (thread == row)
    const int r = tid;
    const int64_t L = (int64_t)((int)blockIdx.x - cl.tile0) * 128 + r;
    const bool row_ok = L < cl.total_rows;
    int n = 0, rz = 0, rx = 0, ry = 0;
    if (row_ok) {
      int64_t q = L;
      ry = (int)(q % cl.RY); q /= cl.RY;
      rx = (int)(q % cl.RX); q /= cl.RX;
      rz = (int)(q % cl.RZ);
      n = (int)(q / cl.RZ);
    }
    const bf16* in_row = p.in.ptr + (int64_t)n * p.in.n_stride + p.in.voxel(rz * p.IS, rx * p.IS, ry * p.IS);
    const int izc = rz * p.IS;

    auto issue = [&](int item) {
      const int st = item % kStages;
      const int tap = item / p.nk, kc = item - tap * p.nk;
      const int c0 = kc * kch;
      uint8_t* sA = smem + (size_t)st * stage_bytes;
      uint8_t* sB = sA + a_bytes;
      const int tz = cl.tz[tap], tx = cl.tx[tap], ty = cl.ty[tap];
      const int iz = izc + tz;
      const bool ok = row_ok && iz >= 0 && iz < p.in.Z;
      const bf16* src = in_row + (((int64_t)tz * p.in.Xp + tx) * p.in.Yp + ty) * 8;
      for (int c = 0; c < kch; ++c) {
        const bool okc = ok && (c0 + c) < p.in.C8;
        cp_async16(sA + ((size_t)c * 128 + r) * 16, okc ? (const void*)(src + (int64_t)(c0 + c) * p.in.c_stride) : (const void*)p.in.ptr, okc);
      }
      const bf16* wsrc = p.w + (size_t)cl.wt[tap] * p.Cin * p.NP + (size_t)kc * p.KC * p.NP;
      const int nb16 = (p.KC * p.NP) >> 3;
      for (int i = r; i < nb16; i += 128) cp_async16(sB + (size_t)i * 16, wsrc + (size_t)i * 8, true);
      cp_async_commit();
    };
    auto publish = [&](int item) {  // this thread's cp.async group for `item` has landed
      fence_proxy_async();
      mbar_arrive(&bar_full[item % kStages]);
    };

    for (int item = 0; item < nitems; ++item) {
      if (item >= kStages) mbar_wait(&bar_empty[item % kStages], ((item / kStages) - 1) & 1);
      issue(item);
      if (item >= kStages - 1) {
        cp_async_wait<kStages - 1>();
        publish(item - (kStages - 1));
      }
    }
    // drain the last (up to kStages-1) groups
    cp_async_wait<0>();
    for (int item = (nitems >= kStages - 1 ? nitems - (kStages - 1) : 0); item < nitems; ++item) publish(item);

    // ------------------------------------------------------------------ epilogue
    mbar_wait(&bar_acc, 0);
    fence_after_sync();
    const int oz = rz * p.OS + cl.oz0, ox = rx * p.OS + cl.ox0, oy = ry * p.OS + cl.oy0;
    const int64_t ovox = p.out.voxel(oz, ox, oy);
    bf16* out_row = p.out.ptr + (int64_t)n * p.out.n_stride + ovox;
    const bf16* res_row = p.has_res ? p.res.ptr + (int64_t)n * p.res.n_stride + p.res.voxel(oz, ox, oy) : nullptr;
    const bf16* mask_row = p.has_mask ? p.mask.ptr + (int64_t)n * p.mask.n_stride + p.mask.voxel(oz, ox, oy) : nullptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c16 = 0; c16 * 16 < p.NP; ++c16) {
      uint32_t v[16];
      tmem_ld16(trow + c16 * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ch = c16 * 2 + h;
        if (ch >= p.out_c8 || !row_ok) continue;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[h * 8 + i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += __ldg(p.bias + ch * 8 + i);
        }
        if (res_row) {
          float g[8];
          unpack8(ldg16(res_row + ch * p.res.c_stride), g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += g[i];
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (mask_row) {
          float g[8];
          unpack8(ldg16(mask_row + ch * p.mask.c_stride), g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = g[i] > 0.f ? f[i] : 0.f;
        }
        bf16* dst = out_row + ch * p.out.c_stride;
        if (p.accumulate) {
          float g[8];
          unpack8(*reinterpret_cast<const uint4*>(dst), g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += g[i];
        }
        stg16(dst, pack8(f));
      }
    }
  } else if (tid == 128) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = idesc_bf16(128, p.NP, 0, 0);
    for (int item = 0; item < nitems; ++item) {
      const int st = item % kStages;
      mbar_wait(&bar_full[st], (item / kStages) & 1);
      fence_after_sync();
      const uint32_t sA = smem_u32(smem + (size_t)st * stage_bytes);
      const uint32_t sB = sA + a_bytes;
      for (int k16 = 0; k16 < (p.KC >> 4); ++k16) {
        const uint64_t ad = smem_desc(sA + k16 * 2 * 2048, 2048, 128);
        const uint64_t bd = smem_desc(sB + k16 * 2 * p.NP * 16, p.NP * 16, 128);
        mma_ss(tmem, ad, bd, idesc, (item | k16) ? 1u : 0u);
      }
      mma_commit(&bar_empty[st]);
    }
    mma_commit(&bar_acc);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
  }
}

int check_view(const rtp_p8& t, const char* what) {
  RTP_CHECK_ARG(t.ptr != nullptr, "rtp_conv: %s is null", what);
  RTP_CHECK_ARG(((uintptr_t)t.ptr & 15) == 0 && t.n_stride % 8 == 0 && t.c_stride % 8 == 0, "rtp_conv: %s misaligned", what);
  return 0;
}

}  // namespace

namespace {
int check_class(const rtp_conv_desc* d) {
  RTP_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= RTP_MAX_TAPS, "rtp_conv: bad ntaps=%d", d->ntaps);
  RTP_CHECK_ARG(d->RZ > 0 && d->RX > 0 && d->RY > 0, "rtp_conv: empty row grid");
  // the row grid must stay inside both volumes (in-plane taps may touch the zero pad ring only)
  RTP_CHECK_ARG((d->RZ - 1) * d->OS + d->oz0 < d->out.Z && (d->RX - 1) * d->OS + d->ox0 < d->out.X &&
                    (d->RY - 1) * d->OS + d->oy0 < d->out.Y,
                "rtp_conv: row grid exceeds the output volume");
  for (int t = 0; t < d->ntaps; ++t) {
    RTP_CHECK_ARG(d->tx[t] >= -1 && (d->RX - 1) * d->IS + d->tx[t] <= d->in.X && d->ty[t] >= -1 &&
                      (d->RY - 1) * d->IS + d->ty[t] <= d->in.Y,
                  "rtp_conv: tap %d leaves the padded input plane", t);
  }
  return 0;
}

// descs[0..n): same tensors / weights / strides / epilogue flags, different tap lists and row grids
int conv_launch(const rtp_conv_desc* descs, int n, void* stream) {
  const rtp_conv_desc* d = descs;
  RTP_CHECK_ARG(d != nullptr && n >= 1 && n <= kMaxClasses, "rtp_conv: bad descriptor list");
  if (check_view(d->in, "in") || check_view(d->out, "out")) return -1;
  RTP_CHECK_ARG(d->w != nullptr, "rtp_conv: null weights");
  RTP_CHECK_ARG(d->Cin >= 16 && d->Cin % 16 == 0 && d->Cin <= 512, "rtp_conv: Cin=%d must be a multiple of 16", d->Cin);
  RTP_CHECK_ARG(d->NP >= 16 && d->NP % 16 == 0 && d->NP <= 256, "rtp_conv: NP=%d must be a multiple of 16 <= 256", d->NP);
  RTP_CHECK_ARG(d->out_c8 >= 1 && d->out_c8 * 8 <= d->NP && d->out_c8 <= d->out.C8, "rtp_conv: bad out_c8=%d", d->out_c8);
  RTP_CHECK_ARG((d->IS == 1 || d->IS == 2) && (d->OS == 1 || d->OS == 2), "rtp_conv: strides must be 1 or 2");
  RTP_CHECK_ARG(d->in.N == d->out.N, "rtp_conv: batch mismatch");
  if (d->res.ptr && check_view(d->res, "res")) return -1;
  if (d->mask.ptr && check_view(d->mask, "mask")) return -1;

  ConvK k;
  k.in = P8(d->in); k.out = P8(d->out); k.res = P8(d->res); k.mask = P8(d->mask);
  k.w = (const bf16*)d->w; k.bias = d->bias;
  k.Cin = d->Cin; k.NP = d->NP; k.out_c8 = d->out_c8;
  k.IS = d->IS; k.OS = d->OS; k.relu = d->relu; k.accumulate = d->accumulate;
  k.tmem_cols = 32;
  while (k.tmem_cols < d->NP) k.tmem_cols *= 2;
  k.has_res = d->res.ptr != nullptr; k.has_mask = d->mask.ptr != nullptr;
  k.ncls = n;
  int64_t tiles = 0;
  for (int c = 0; c < n; ++c) {
    const rtp_conv_desc* e = descs + c;
    RTP_CHECK_ARG(e->in.ptr == d->in.ptr && e->out.ptr == d->out.ptr && e->res.ptr == d->res.ptr && e->mask.ptr == d->mask.ptr &&
                      e->w == d->w && e->bias == d->bias && e->Cin == d->Cin && e->NP == d->NP && e->out_c8 == d->out_c8 &&
                      e->IS == d->IS && e->OS == d->OS && e->relu == d->relu && e->accumulate == d->accumulate,
                  "rtp_conv_multi: descriptor %d differs in more than taps / row grid", c);
    if (check_class(e)) return -1;
    ConvClass& cl = k.cls[c];
    for (int t = 0; t < RTP_MAX_TAPS; ++t) { cl.tz[t] = e->tz[t]; cl.tx[t] = e->tx[t]; cl.ty[t] = e->ty[t]; cl.wt[t] = e->wt[t]; }
    cl.ntaps = e->ntaps; cl.RZ = e->RZ; cl.RX = e->RX; cl.RY = e->RY; cl.oz0 = e->oz0; cl.ox0 = e->ox0; cl.oy0 = e->oy0;
    cl.total_rows = (int64_t)e->in.N * e->RZ * e->RX * e->RY;
    cl.tile0 = (int)tiles;
    tiles += (cl.total_rows + 127) / 128;
  }
  RTP_CHECK_ARG(tiles < (1ll << 31), "rtp_conv: too many tiles");

  k.KC = d->Cin; k.nk = 1;
  while (((size_t)k.KC * 256 + (size_t)k.KC * d->NP * 2) * 2 > 200 * 1024 && k.KC % 32 == 0) { k.KC /= 2; k.nk *= 2; }
  const size_t stage = (size_t)k.KC * 256 + (size_t)k.KC * d->NP * 2;
  const int stages = 4 * stage <= 200 * 1024 ? 4 : (3 * stage <= 200 * 1024 ? 3 : 2);
  const size_t smem = stages * stage;
  RTP_CHECK_ARG(smem <= 200 * 1024, "rtp_conv: Cin=%d NP=%d needs %zu B of shared memory", d->Cin, d->NP, smem);
  auto kern = stages == 4 ? conv_generic_kernel<4> : (stages == 3 ? conv_generic_kernel<3> : conv_generic_kernel<2>);
  static size_t configured_dev[RTP_MAX_DEVICES][5];  /* the opt-in is per device */
  size_t* configured = configured_dev[rtp_current_device()];
  if (smem > configured[stages]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured[stages] = smem;
  }
  kern<<<(unsigned)tiles, kThreads, smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}
}  // namespace

extern "C" int rtp_conv(const rtp_conv_desc* d, void* stream) { return conv_launch(d, 1, stream); }

extern "C" int rtp_conv_multi(const rtp_conv_desc* descs, int32_t n, void* stream) { return conv_launch(descs, n, stream); }
