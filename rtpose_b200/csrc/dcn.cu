// dcn.cu — deformable convolution v1 and v2 ("modulated": a per-tap mask multiplies every sample, optional bias)
// (2-D, groups = 1) forward / backward without the reference's `columns` round trip through HBM
// (det3d/ops/dcn/src/deform_conv_cuda.cpp:196-247 materialises C*kh*kw x N*Ho*Wo fp32; the modulated path :490-684).
// Every kernel takes an optional `mask` [N][dg*kh*kw][Ho][Wo]; nullptr = v1.
//
// fp32 kernels (the default; the reference op's dtype and NCHW layout in and out): each CTA owns a tile of output pixels
// of one sample (64 in the forward, 32 in the two backward kernels).  A channel tile of the deformed im2col matrix is
// sampled straight into shared memory with warp-level bilinear gathers (lanes = consecutive output pixels, so the offset
// reads and most of the 4-corner reads coalesce) and is consumed in place by the contraction with the weights on the fp32
// CUDA cores.  Sampling semantics follow deform_conv_cuda_kernel.cu:84-115 (bilinear, zero outside) and :229 (validity
// h>-1 && w>-1 && h<H && w<W); offset channel order [dg][kh*kw][dy,dx].
//
// Tensor-core path (opt-in from rtpose_b200/dcn.py): rtp_dcn_sample_p8 writes the sample volume in bf16 P8 with the taps
// on the z axis, rtp_conv / rtp_wgrad contract it on tcgen05, rtp_dcn_col2im_p8 scatters the sample gradient.
#include "common.cuh"

namespace {

constexpr int kPix = 32;   // output pixels per CTA (backward kernels)
constexpr int kCT = 8;     // input channels per smem tile

struct Dcn {
  int N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, Ho, Wo;
};

struct Sample {
  float w1, w2, w3, w4;  // corner weights (0 when the corner is outside)
  int o1, o2, o3, o4;    // corner offsets inside one channel plane
  float lh, lw;
  bool valid;
};

__device__ __forceinline__ Sample make_sample(const Dcn& p, float h, float w) {
  Sample s;
  s.valid = h > -1.f && w > -1.f && h < (float)p.H && w < (float)p.W;
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  s.lh = h - (float)h_low;
  s.lw = w - (float)w_low;
  const float hh = 1.f - s.lh, hw = 1.f - s.lw;
  // corner flags include `valid`: an invalid position (also NaN / far outside) must never turn into a load address
  const bool t = s.valid && h_low >= 0, b = s.valid && h_high <= p.H - 1, l = s.valid && w_low >= 0, r = s.valid && w_high <= p.W - 1;
  s.w1 = (t && l) ? hh * hw : 0.f;
  s.w2 = (t && r) ? hh * s.lw : 0.f;
  s.w3 = (b && l) ? s.lh * hw : 0.f;
  s.w4 = (b && r) ? s.lh * s.lw : 0.f;
  s.o1 = (t && l) ? h_low * p.W + w_low : 0;
  s.o2 = (t && r) ? h_low * p.W + w_high : 0;
  s.o3 = (b && l) ? h_high * p.W + w_low : 0;
  s.o4 = (b && r) ? h_high * p.W + w_high : 0;
  return s;
}

// sampling position of tap (i, j) at output pixel (ho, wo) for deformable group g
__device__ __forceinline__ void tap_pos(const Dcn& p, const float* __restrict__ off_n, int g, int t, int ho, int wo, float& h,
                                        float& w) {
  const int i = t / p.kw, j = t - i * p.kw;
  const int K = p.kh * p.kw;
  const int64_t plane = (int64_t)p.Ho * p.Wo;
  const float* o = off_n + ((int64_t)(g * K + t) * 2) * plane + (int64_t)ho * p.Wo + wo;
  h = (float)(ho * p.stride - p.pad + i * p.dil) + __ldg(o);
  w = (float)(wo * p.stride - p.pad + j * p.dil) + __ldg(o + plane);
}

// col[ck][px] for channels [c0, c0+kCT): one warp-level gather per (channel, tap) row
template <int PIX = kPix>
__device__ __forceinline__ void sample_tile(const Dcn& p, const float* __restrict__ x_n, const float* __restrict__ off_n,
                                            const float* __restrict__ mask_n, int c0, int pix0, float* col) {
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo, cpg = p.C / p.dg;
  for (int i = threadIdx.x; i < kCT * K * PIX; i += blockDim.x) {
    const int px = i % PIX, ck = i / PIX;
    const int c = c0 + ck / K, t = ck % K;
    float v = 0.f;
    const int pix = pix0 + px;
    if (c < p.C && pix < npix) {
      const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
      float h, w;
      tap_pos(p, off_n, c / cpg, t, ho, wo, h, w);
      const Sample s = make_sample(p, h, w);
      const float* xc = x_n + (int64_t)c * p.H * p.W;
      v = s.w1 * __ldg(xc + s.o1) + s.w2 * __ldg(xc + s.o2) + s.w3 * __ldg(xc + s.o3) + s.w4 * __ldg(xc + s.o4);
      if (mask_n) v *= __ldg(mask_n + (int64_t)((c / cpg) * K + t) * npix + pix);  // modulated_deformable_im2col (kernel.cu:571-634)
    }
    col[ck * PIX + px] = v;
  }
}

// y[n, co, pix] = bias[co] + sum_{c,t} w[co, c, t] * col[(c,t), pix]
// CTA tile: 64 pixels x 128 output channels, 256 threads, 4 pixels x 8 channels of accumulators per thread.  Per channel
// tile (8 channels x K taps) the deformed samples AND the matching weight slice are staged in shared memory, so the inner
// loop is 3 LDS.128 (4 pixels, 8 weights broadcast within a half-warp) per 32 FMAs.  The weight slice is read from the
// reference's [Cout][C][kh][kw] layout in 32-byte runs along (c, tap) — a warp covers 4 output channels x 8 consecutive
// (c, tap) — and stored transposed with a row pitch of 132 floats: bank = (4*ck + co) mod 32 is distinct for those 32 lanes.
constexpr int kFPix = 64, kFCo = 128, kWPitch = kFCo + 4;
__global__ void __launch_bounds__(256) dcn_fwd_kernel(Dcn p, const float* __restrict__ x, const float* __restrict__ off,
                                                      const float* __restrict__ mask, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y) {
  extern __shared__ float4 dcn_smem4[];
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo, nckmax = kCT * K;
  float* col = reinterpret_cast<float*>(dcn_smem4);  // [kCT*K][kFPix]
  float* wsm = col + nckmax * kFPix;                 // [kCT*K][kWPitch]
  const int n = blockIdx.y, pix0 = blockIdx.x * kFPix, co0 = blockIdx.z * kFCo;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // pixels pix0 + 4*tx + {0..3}, channels co0 + 8*ty + {0..7}
  const float* x_n = x + (int64_t)n * p.C * p.H * p.W;
  const float* off_n = off + (int64_t)n * p.dg * K * 2 * npix;
  const float* mask_n = mask ? mask + (int64_t)n * p.dg * K * npix : nullptr;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = co0 + ty * 8 + j;
    const float b = (bias && co < p.Cout) ? __ldg(bias + co) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][j] = b;
  }
  // The sampling geometry (4 corner weights incl. the mask, 4 corner offsets) depends on (deformable group, tap, pixel)
  // only: when channel tiles do not straddle groups it is computed once per group into shared memory and every channel
  // of the group then costs 2 LDS.128 + 4 independent gathers, with no offset-load -> address -> gather chain.
  float4* geo = reinterpret_cast<float4*>(col + ((nckmax * (kFPix + kWPitch) + 3) & ~3));  // [K*kFPix][2]
  const int cpg = p.C / p.dg;
  const bool cached = (cpg % kCT) == 0;
  int cur_g = -1;
  for (int c0 = 0; c0 < p.C; c0 += kCT) {
    __syncthreads();
    // weight slice of this channel tile: asynchronous 4-byte copies (zero-filled beyond Cout), issued first so that their
    // L2 latency overlaps the gathers below
    const int nck = min(kCT, p.C - c0) * K;
    {
      const int nckb = (nck + 7) >> 3;
      const uint32_t wsm_s = (uint32_t)__cvta_generic_to_shared(wsm);
      for (int i = threadIdx.x; i < (kFCo / 4) * nckb * 32; i += blockDim.x) {
        const int q = i >> 5, cob = q / nckb, ckb = q - cob * nckb;
        const int ck = ckb * 8 + (i & 7), co = cob * 4 + ((i >> 3) & 3);
        if (ck < nck) {
          const int cg = min(co0 + co, p.Cout - 1);
          const float* src = w + ((int64_t)cg * p.C + c0) * K + ck;
          const int nbytes = (co0 + co < p.Cout) ? 4 : 0;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(wsm_s + (uint32_t)(ck * kWPitch + co) * 4u), "l"(src), "r"(nbytes)
                       : "memory");
        }
      }
    }
    if (cached) {
      const int g = c0 / cpg;
      if (g != cur_g) {  // uniform over the CTA
        for (int i = threadIdx.x; i < K * kFPix; i += blockDim.x) {
          const int t = i / kFPix, pix = pix0 + (i - t * kFPix);
          float4 wq = make_float4(0.f, 0.f, 0.f, 0.f);
          int4 oq = make_int4(0, 0, 0, 0);
          if (pix < npix) {
            const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
            float h, wv;
            tap_pos(p, off_n, g, t, ho, wo, h, wv);
            const Sample sm = make_sample(p, h, wv);
            const float mk = mask_n ? __ldg(mask_n + (int64_t)(g * K + t) * npix + pix) : 1.f;
            wq = make_float4(sm.w1 * mk, sm.w2 * mk, sm.w3 * mk, sm.w4 * mk);
            oq = make_int4(sm.o1, sm.o2, sm.o3, sm.o4);
          }
          geo[2 * i] = wq;
          geo[2 * i + 1] = *reinterpret_cast<float4*>(&oq);
        }
        cur_g = g;
        __syncthreads();
      }
      const int64_t plane = (int64_t)p.H * p.W;
#pragma unroll 6
      for (int i = threadIdx.x; i < kCT * K * kFPix; i += blockDim.x) {
        const int px = i % kFPix, ck = i / kFPix;
        const int cc = ck / K, t = ck - cc * K;
        const float4 wq = geo[2 * (t * kFPix + px)];
        const float4 of = geo[2 * (t * kFPix + px) + 1];
        const int4 oq = *reinterpret_cast<const int4*>(&of);
        const float* xc = x_n + (int64_t)min(c0 + cc, p.C - 1) * plane;
        col[ck * kFPix + px] = wq.x * __ldg(xc + oq.x) + wq.y * __ldg(xc + oq.y) + wq.z * __ldg(xc + oq.z) + wq.w * __ldg(xc + oq.w);
      }
    } else {
      sample_tile<kFPix>(p, x_n, off_n, mask_n, c0, pix0, col);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    for (int ck = 0; ck < nck; ++ck) {
      const float4 v = *reinterpret_cast<const float4*>(col + ck * kFPix + tx * 4);
      const float4 wa = *reinterpret_cast<const float4*>(wsm + ck * kWPitch + ty * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wsm + ck * kWPitch + ty * 8 + 4);
      const float wr[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float wv = wr[j];
        acc[0][j] = fmaf(wv, v.x, acc[0][j]);
        acc[1][j] = fmaf(wv, v.y, acc[1][j]);
        acc[2][j] = fmaf(wv, v.z, acc[2][j]);
        acc[3][j] = fmaf(wv, v.w, acc[3][j]);
      }
    }
  }
  const int pix = pix0 + tx * 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = co0 + ty * 8 + j;
    if (co >= p.Cout) continue;
    float* yr = y + ((int64_t)n * p.Cout + co) * npix + pix;
    if ((npix & 3) == 0 && pix + 3 < npix) {
      *reinterpret_cast<float4*>(yr) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (pix + q < npix) yr[q] = acc[q][j];
    }
  }
}

// dx (atomic scatter) and doffset for one pixel tile
__global__ void __launch_bounds__(256) dcn_bwd_input_kernel(Dcn p, const float* __restrict__ x, const float* __restrict__ off,
                                                            const float* __restrict__ mask, const float* __restrict__ w,
                                                            const float* __restrict__ dy, float* __restrict__ dx,
                                                            float* __restrict__ doff, float* __restrict__ dmask) {
  extern __shared__ float sm[];
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo, cpg = p.C / p.dg;
  float* dcol = sm;                      // [kCT*K][kPix]
  float* dys = sm + kCT * K * kPix;      // [Cout][kPix]
  const int n = blockIdx.y, pix0 = blockIdx.x * kPix;
  const float* x_n = x + (int64_t)n * p.C * p.H * p.W;
  const float* off_n = off + (int64_t)n * p.dg * K * 2 * npix;
  float* dx_n = dx + (int64_t)n * p.C * p.H * p.W;
  float* doff_n = doff + (int64_t)n * p.dg * K * 2 * npix;
  const float* mask_n = mask ? mask + (int64_t)n * p.dg * K * npix : nullptr;
  float* dmask_n = mask ? dmask + (int64_t)n * p.dg * K * npix : nullptr;
  for (int i = threadIdx.x; i < p.Cout * kPix; i += blockDim.x) {
    const int px = i % kPix, co = i / kPix;
    dys[i] = (pix0 + px < npix) ? dy[((int64_t)n * p.Cout + co) * npix + pix0 + px] : 0.f;
  }
  // each thread owns fixed (tap, pixel) pairs across all channel tiles, so its doffset sums need no atomics
  for (int c0 = 0; c0 < p.C; c0 += kCT) {
    __syncthreads();
    // dcol[(c,t), px] = sum_co w[co, c, t] * dy[co, px]
    for (int i = threadIdx.x; i < kCT * K * kPix; i += blockDim.x) {
      const int px = i % kPix, ck = i / kPix;
      float a = 0.f;
      if (c0 + ck / K < p.C) {
        const float* wp = w + (int64_t)c0 * K + ck;
        for (int co = 0; co < p.Cout; ++co) a = fmaf(__ldg(wp + (int64_t)co * p.C * K), dys[co * kPix + px], a);
      }
      dcol[i] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * kPix; i += blockDim.x) {
      const int px = i % kPix, t = i / kPix;
      const int pix = pix0 + px;
      if (pix >= npix) continue;
      const int ho = pix / p.Wo, wo = pix - ho * p.Wo;
      for (int cc = 0; cc < kCT && c0 + cc < p.C; ++cc) {
        const int c = c0 + cc, g = c / cpg;
        float h, wv;
        tap_pos(p, off_n, g, t, ho, wo, h, wv);
        const Sample s = make_sample(p, h, wv);
        const float draw = dcol[(cc * K + t) * kPix + px];  // gradient w.r.t. the masked sample
        const float mk = mask_n ? __ldg(mask_n + (int64_t)(g * K + t) * npix + pix) : 1.f;
        const float d = draw * mk;                          // gradient w.r.t. the bilinear sample
        const float* xc = x_n + (int64_t)c * p.H * p.W;
        float* dxc = dx_n + (int64_t)c * p.H * p.W;
        if (s.w1 != 0.f) atomicAdd(dxc + s.o1, d * s.w1);
        if (s.w2 != 0.f) atomicAdd(dxc + s.o2, d * s.w2);
        if (s.w3 != 0.f) atomicAdd(dxc + s.o3, d * s.w3);
        if (s.w4 != 0.f) atomicAdd(dxc + s.o4, d * s.w4);
        if (s.valid) {
          // d val / d h and d val / d w of the bilinear form (corners outside the image contribute 0)
          const int h_low = (int)floorf(h), w_low = (int)floorf(wv);
          const bool tt = h_low >= 0, bb = h_low + 1 <= p.H - 1, ll = w_low >= 0, rr = w_low + 1 <= p.W - 1;
          const float x1 = (tt && ll) ? __ldg(xc + s.o1) : 0.f, x2 = (tt && rr) ? __ldg(xc + s.o2) : 0.f;
          const float x3 = (bb && ll) ? __ldg(xc + s.o3) : 0.f, x4 = (bb && rr) ? __ldg(xc + s.o4) : 0.f;
          const float hw = 1.f - s.lw, hh = 1.f - s.lh;
          const float gh = -hw * x1 - s.lw * x2 + hw * x3 + s.lw * x4;
          const float gw = -hh * x1 + hh * x2 - s.lh * x3 + s.lh * x4;
          const int64_t plane = (int64_t)npix;
          float* o = doff_n + ((int64_t)(g * K + t) * 2) * plane + pix;
          o[0] += d * gh;      // this thread is the only writer of (n, g, t, pix) — channels are visited sequentially
          o[plane] += d * gw;
          // d / d mask = the unmodulated sample (modulated_deformable_col2im_coord, kernel.cu:696-767)
          if (dmask_n) dmask_n[(int64_t)(g * K + t) * plane + pix] += draw * (s.w1 * x1 + s.w2 * x2 + s.w3 * x3 + s.w4 * x4);
        }
      }
    }
  }
}

// dw[co, c, t] += scale * sum_{pixels of this chunk} dy[co, pix] * col[(c,t), pix]
__global__ void __launch_bounds__(256) dcn_bwd_weight_kernel(Dcn p, const float* __restrict__ x, const float* __restrict__ off,
                                                             const float* __restrict__ mask, const float* __restrict__ dy,
                                                             float* __restrict__ dw, float scale, int tiles_per_block) {
  extern __shared__ float sm[];
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo;
  float* col = sm;                   // [kCT*K][kPix]
  float* dys = sm + kCT * K * kPix;  // [kPix][64]  (transposed: conflict-free for consecutive co)
  const int n = blockIdx.y, c0 = blockIdx.z * kCT;
  const int co_l = threadIdx.x & 63, ckg = threadIdx.x >> 6;  // 4 groups of ck
  const int nck = kCT * K, per = (nck + 3) / 4;
  const float* x_n = x + (int64_t)n * p.C * p.H * p.W;
  const float* off_n = off + (int64_t)n * p.dg * K * 2 * npix;
  const float* mask_n = mask ? mask + (int64_t)n * p.dg * K * npix : nullptr;
  for (int cob = 0; cob < p.Cout; cob += 64) {
    float acc[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) acc[j] = 0.f;
    for (int tl = 0; tl < tiles_per_block; ++tl) {
      const int pix0 = (blockIdx.x * tiles_per_block + tl) * kPix;
      if (pix0 >= npix) break;
      __syncthreads();
      sample_tile(p, x_n, off_n, mask_n, c0, pix0, col);
      for (int i = threadIdx.x; i < 64 * kPix; i += blockDim.x) {
        const int px = i % kPix, co = i / kPix;
        dys[px * 64 + co] = (cob + co < p.Cout && pix0 + px < npix) ? dy[((int64_t)n * p.Cout + cob + co) * npix + pix0 + px] : 0.f;
      }
      __syncthreads();
      for (int px = 0; px < kPix; ++px) {
        const float a = dys[px * 64 + co_l];
#pragma unroll
        for (int j = 0; j < 18; ++j) {
          const int ck = ckg * per + j;
          if (j < per && ck < nck) acc[j] = fmaf(a, col[ck * kPix + px], acc[j]);
        }
      }
    }
    const int co = cob + co_l;
    if (co < p.Cout) {
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        const int ck = ckg * per + j;
        if (j < per && ck < nck && c0 + ck / K < p.C) atomicAdd(dw + ((int64_t)co * p.C + c0) * K + ck, scale * acc[j]);
      }
    }
  }
}

// Tensor-core path, step 1: the deformed im2col tensor as a bf16 P8 volume whose z axis is the tap index,
// S[n][c/8][t][wo][ho][c%8] = mask * bilinear(x[n, c], p(ho, wo) + tap t + offset); step 2 is rtp_conv on it with the tap
// list {(tz = t, 0, 0)} — a 1x1 conv over K = C*kh*kw accumulated in TMEM (fp32).
// One thread per (n, deformable group, tap, ho, wo) computes the sampling geometry once and reuses it for all channels
// of the group.  A CTA covers 32 wo x 8 ho: while sampling, lanes run along wo (the fastest axis of x and of the offsets,
// so the corner reads of neighbouring lanes share cache lines); each 8-channel result goes through a shared-memory
// transpose so that the stores run along ho, the fastest P8 axis (8 x 16 B = 128 contiguous bytes per wo).
__global__ void __launch_bounds__(256) dcn_sample_p8_kernel(Dcn p, const float* __restrict__ x, const float* __restrict__ off,
                                                            const float* __restrict__ mask, P8 dst) {
  __shared__ uint4 stage[8][33];
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo, cpg = p.C / p.dg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wo = blockIdx.x * 32 + lane, ho = blockIdx.y * 8 + warp;
  int b = blockIdx.z;
  const int t = b % K;
  b /= K;
  const int g = b % p.dg, n = b / p.dg;
  const bool live = ho < p.Ho && wo < p.Wo;
  Sample s;
  s.w1 = s.w2 = s.w3 = s.w4 = 0.f;
  s.o1 = s.o2 = s.o3 = s.o4 = 0;
  if (live) {
    const float* off_n = off + (int64_t)n * p.dg * K * 2 * npix;
    float h, w;
    tap_pos(p, off_n, g, t, ho, wo, h, w);
    s = make_sample(p, h, w);
    if (mask) {
      const float m = __ldg(mask + ((int64_t)n * p.dg * K + g * K + t) * npix + (int64_t)ho * p.Wo + wo);
      s.w1 *= m, s.w2 *= m, s.w3 *= m, s.w4 *= m;
    }
  }
  const int64_t plane = (int64_t)p.H * p.W;
  const float* xg = x + ((int64_t)n * p.C + (int64_t)g * cpg) * plane;
  // store role: 8 consecutive threads cover the 8 ho of one wo
  const int s_ho = threadIdx.x & 7, s_wo = threadIdx.x >> 3;
  const int st_ho = blockIdx.y * 8 + s_ho, st_wo = blockIdx.x * 32 + s_wo;
  const bool st_live = st_ho < p.Ho && st_wo < p.Wo;
  bf16* d = dst.ptr + (int64_t)n * dst.n_stride + dst.voxel(t, st_wo, st_ho);
  for (int c0 = 0; c0 < cpg; c0 += 8) {  // cpg % 8 == 0 (checked by the host): a chunk never straddles two groups
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* xc = xg + (int64_t)(c0 + k) * plane;
      v[k] = s.w1 * __ldg(xc + s.o1) + s.w2 * __ldg(xc + s.o2) + s.w3 * __ldg(xc + s.o3) + s.w4 * __ldg(xc + s.o4);
    }
    stage[warp][lane] = pack8(v);
    __syncthreads();
    if (st_live) stg16(d + (int64_t)((g * cpg + c0) >> 3) * dst.c_stride, stage[s_ho][s_wo]);
    __syncthreads();
  }
}

// Tensor-core backward, last step: dS (bf16 P8, taps on z; produced by rtp_conv from dy) -> dx (atomic scatter through
// the bilinear weights), doffset and dmask.  One thread per (n, deformable group, tap, ho, wo) walks the channels of its
// group, so doffset / dmask are plain stores of register sums.  Lanes run along wo — the contiguous axis of x, dx, offset
// and doffset (NCHW) — so a warp's four corner gathers and four atomic adds per channel fall into one or two 128-byte lines
// instead of 32 (with lanes along ho, dS's contiguous axis, this kernel was 66 % of the deformable head's step: 7 ms per
// launch at C = 256); the price is a 16-byte-per-lane strided read of dS, 1/64 of the traffic.
__global__ void __launch_bounds__(256, 2) dcn_col2im_p8_kernel(Dcn p, const float* __restrict__ x, const float* __restrict__ off,
                                                            const float* __restrict__ mask, P8 ds, float* __restrict__ dx,
                                                            float* __restrict__ doff, float* __restrict__ dmask) {
  const int K = p.kh * p.kw, npix = p.Ho * p.Wo, cpg = p.C / p.dg;
  const int wo = blockIdx.x * 32 + (threadIdx.x & 31), ho = blockIdx.y * 8 + (threadIdx.x >> 5);
  int b = blockIdx.z;
  const int t = b % K;
  b /= K;
  const int g = b % p.dg, n = b / p.dg;
  if (ho >= p.Ho || wo >= p.Wo) return;
  const int pix = ho * p.Wo + wo;
  const float* off_n = off + (int64_t)n * p.dg * K * 2 * npix;
  float h, w;
  tap_pos(p, off_n, g, t, ho, wo, h, w);
  const Sample s = make_sample(p, h, w);
  const float mk = mask ? __ldg(mask + ((int64_t)n * p.dg * K + g * K + t) * npix + pix) : 1.f;
  const int64_t plane = (int64_t)p.H * p.W;
  const float* xg = x + ((int64_t)n * p.C + (int64_t)g * cpg) * plane;
  float* dxg = dx + ((int64_t)n * p.C + (int64_t)g * cpg) * plane;
  const bf16* d = ds.ptr + (int64_t)n * ds.n_stride + ds.voxel(t, wo, ho);
  const float hw = 1.f - s.lw, hh = 1.f - s.lh;
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const bool tt = s.valid && h_low >= 0, bb = s.valid && h_low + 1 <= p.H - 1, ll = s.valid && w_low >= 0, rr = s.valid && w_low + 1 <= p.W - 1;
  float goh = 0.f, gow = 0.f, gm = 0.f;
  for (int c0 = 0; c0 < cpg; c0 += 8) {
    float dv[8];
    unpack8(ldg16(d + (int64_t)((g * cpg + c0) >> 3) * ds.c_stride), dv);
    if (!s.valid) continue;  // an invalid position contributes nothing to any gradient
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* xc = xg + (int64_t)(c0 + k) * plane;
      float* dxc = dxg + (int64_t)(c0 + k) * plane;
      const float draw = dv[k], dd = draw * mk;
      // a corner outside the image reads as 0 in the slopes even where its bilinear weight happens to be 0
      const float x1 = (tt && ll) ? __ldg(xc + s.o1) : 0.f, x2 = (tt && rr) ? __ldg(xc + s.o2) : 0.f;
      const float x3 = (bb && ll) ? __ldg(xc + s.o3) : 0.f, x4 = (bb && rr) ? __ldg(xc + s.o4) : 0.f;
      if (s.w1 != 0.f) atomicAdd(dxc + s.o1, dd * s.w1);
      if (s.w2 != 0.f) atomicAdd(dxc + s.o2, dd * s.w2);
      if (s.w3 != 0.f) atomicAdd(dxc + s.o3, dd * s.w3);
      if (s.w4 != 0.f) atomicAdd(dxc + s.o4, dd * s.w4);
      goh += dd * (-hw * x1 - s.lw * x2 + hw * x3 + s.lw * x4);
      gow += dd * (-hh * x1 + hh * x2 - s.lh * x3 + s.lh * x4);
      gm += draw * (s.w1 * x1 + s.w2 * x2 + s.w3 * x3 + s.w4 * x4);
    }
  }
  float* o = doff + ((int64_t)n * p.dg * K + g * K + t) * 2 * npix + pix;
  o[0] = goh;
  o[npix] = gow;
  if (dmask) dmask[((int64_t)n * p.dg * K + g * K + t) * npix + pix] = gm;
}

int make(Dcn& d, int N, int C, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil, int dg, const char* who) {
  RTP_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && dil > 0 && dg > 0, "%s: bad sizes", who);
  RTP_CHECK_ARG(C % dg == 0, "%s: input channels %d not divisible by deformable groups %d", who, C, dg);
  RTP_CHECK_ARG(kh * kw * kCT <= 18 * 4, "%s: kernel %dx%d too large (max 9 taps)", who, kh, kw);
  d = Dcn{N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1,
          (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1};
  RTP_CHECK_ARG(d.Ho > 0 && d.Wo > 0, "%s: empty output", who);
  return 0;
}

// dbias[co] += scale * sum_{n, pix} dy[n, co, pix]   (one CTA per output channel)
__global__ void __launch_bounds__(256) dcn_bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ dbias, int N, int Cout,
                                                            int npix, float scale) {
  __shared__ float red[8];
  const int co = blockIdx.x;
  float a = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* r = dy + ((int64_t)n * Cout + co) * npix;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) a += __ldg(r + i);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    dbias[co] += scale * t;
  }
}

int launch_fwd(const Dcn& d, const float* x, const float* offset, const float* mask, const float* w, const float* bias, float* y,
               cudaStream_t st) {
  const size_t smem = (((size_t)kCT * d.kh * d.kw * (kFPix + kWPitch) + 3) & ~(size_t)3) * sizeof(float) +
                      (size_t)d.kh * d.kw * kFPix * 2 * sizeof(float4);
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaFuncSetAttribute(dcn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dim3 grid(ceil_div((int64_t)d.Ho * d.Wo, kFPix), d.N, ceil_div(d.Cout, kFCo));
  dcn_fwd_kernel<<<grid, 256, smem, st>>>(d, x, offset, mask, w, bias, y);
  RTP_LAUNCH_CHECK();
}

int launch_bwd_input(const Dcn& d, const float* x, const float* offset, const float* mask, const float* w, const float* dy, float* dx,
                     float* doffset, float* dmask, cudaStream_t st, const char* who) {
  const size_t smem = ((size_t)kCT * d.kh * d.kw * kPix + (size_t)d.Cout * kPix) * sizeof(float);
  RTP_CHECK_ARG(smem <= 200 * 1024, "%s: Cout=%d too large", who, d.Cout);
  static size_t configured_dev[RTP_MAX_DEVICES];  /* the opt-in is per device */
  size_t& configured = configured_dev[rtp_current_device()];
  if (smem > configured) {
    cudaFuncSetAttribute(dcn_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const size_t taps = (size_t)d.N * d.dg * d.kh * d.kw * d.Ho * d.Wo;
  cudaMemsetAsync(dx, 0, (size_t)d.N * d.C * d.H * d.W * sizeof(float), st);
  cudaMemsetAsync(doffset, 0, taps * 2 * sizeof(float), st);
  if (mask) cudaMemsetAsync(dmask, 0, taps * sizeof(float), st);
  dim3 grid(ceil_div((int64_t)d.Ho * d.Wo, kPix), d.N);
  dcn_bwd_input_kernel<<<grid, 256, smem, st>>>(d, x, offset, mask, w, dy, dx, doffset, dmask);
  RTP_LAUNCH_CHECK();
}

int launch_bwd_weight(const Dcn& d, const float* x, const float* offset, const float* mask, const float* dy, float* dw, float scale,
                      cudaStream_t st) {
  const int tiles = ceil_div((int64_t)d.Ho * d.Wo, kPix);
  const int tiles_per_block = tiles > 64 ? 16 : 1;
  const size_t smem = ((size_t)kCT * d.kh * d.kw * kPix + 64 * kPix) * sizeof(float);
  dim3 grid(ceil_div(tiles, tiles_per_block), d.N, ceil_div(d.C, kCT));
  dcn_bwd_weight_kernel<<<grid, 256, smem, st>>>(d, x, offset, mask, dy, dw, scale, tiles_per_block);
  RTP_LAUNCH_CHECK();
}

}  // namespace

extern "C" int rtp_dcn_fwd(const float* x, const float* offset, const float* w, float* y, int32_t N, int32_t C, int32_t H,
                           int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil,
                           int32_t dg, void* stream) {
  RTP_CHECK_ARG(x && offset && w && y, "rtp_dcn_fwd: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_dcn_fwd")) return -1;
  return launch_fwd(d, x, offset, nullptr, w, nullptr, y, (cudaStream_t)stream);
}

extern "C" int rtp_dcn_bwd_input(const float* x, const float* offset, const float* w, const float* dy, float* dx, float* doffset,
                                 int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw,
                                 int32_t stride, int32_t pad, int32_t dil, int32_t dg, void* stream) {
  RTP_CHECK_ARG(x && offset && w && dy && dx && doffset, "rtp_dcn_bwd_input: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_dcn_bwd_input")) return -1;
  return launch_bwd_input(d, x, offset, nullptr, w, dy, dx, doffset, nullptr, (cudaStream_t)stream, "rtp_dcn_bwd_input");
}

extern "C" int rtp_dcn_bwd_weight(const float* x, const float* offset, const float* dy, float* dw, int32_t N, int32_t C, int32_t H,
                                  int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil,
                                  int32_t dg, float scale, void* stream) {
  RTP_CHECK_ARG(x && offset && dy && dw, "rtp_dcn_bwd_weight: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_dcn_bwd_weight")) return -1;
  return launch_bwd_weight(d, x, offset, nullptr, dy, dw, scale, (cudaStream_t)stream);
}

extern "C" int rtp_dcn_sample_p8(const float* x, const float* offset, const float* mask, rtp_p8 dst, int32_t N, int32_t C, int32_t H,
                                 int32_t W, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil, int32_t dg,
                                 void* stream) {
  RTP_CHECK_ARG(x && offset && dst.ptr, "rtp_dcn_sample_p8: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, 1, kh, kw, stride, pad, dil, dg, "rtp_dcn_sample_p8")) return -1;
  RTP_CHECK_ARG((C / dg) % 8 == 0, "rtp_dcn_sample_p8: channels per deformable group (%d) must be a multiple of 8", C / dg);
  RTP_CHECK_ARG(dst.N == N && dst.C8 * 8 >= C && dst.Z == kh * kw && dst.Y == d.Ho && dst.X == d.Wo,
                "rtp_dcn_sample_p8: dst must be P8 [N=%d][C>=%d][Z=%d taps][Y=%d][X=%d]", N, C, kh * kw, d.Ho, d.Wo);
  RTP_CHECK_ARG((int64_t)N * dg * kh * kw <= 65535, "rtp_dcn_sample_p8: N*dg*taps = %lld exceeds the grid limit; split the batch",
                (long long)N * dg * kh * kw);
  dim3 grid(ceil_div(d.Wo, 32), ceil_div(d.Ho, 8), N * dg * kh * kw);
  dcn_sample_p8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, x, offset, mask, P8(dst));
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_dcn_col2im_p8(const float* x, const float* offset, const float* mask, rtp_p8 ds, float* dx, float* doffset,
                                 float* dmask, int32_t N, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw, int32_t stride,
                                 int32_t pad, int32_t dil, int32_t dg, void* stream) {
  RTP_CHECK_ARG(x && offset && ds.ptr && dx && doffset && (!mask || dmask), "rtp_dcn_col2im_p8: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, 1, kh, kw, stride, pad, dil, dg, "rtp_dcn_col2im_p8")) return -1;
  RTP_CHECK_ARG((C / dg) % 8 == 0, "rtp_dcn_col2im_p8: channels per deformable group (%d) must be a multiple of 8", C / dg);
  RTP_CHECK_ARG(ds.N == N && ds.C8 * 8 >= C && ds.Z == kh * kw && ds.Y == d.Ho && ds.X == d.Wo,
                "rtp_dcn_col2im_p8: ds must be P8 [N=%d][C>=%d][Z=%d taps][Y=%d][X=%d]", N, C, kh * kw, d.Ho, d.Wo);
  RTP_CHECK_ARG((int64_t)N * dg * kh * kw <= 65535, "rtp_dcn_col2im_p8: N*dg*taps exceeds the grid limit; split the batch");
  cudaMemsetAsync(dx, 0, (size_t)N * C * H * W * sizeof(float), (cudaStream_t)stream);
  dim3 grid(ceil_div(d.Wo, 32), ceil_div(d.Ho, 8), N * dg * kh * kw);
  dcn_col2im_p8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, x, offset, mask, P8(ds), dx, doffset, mask ? dmask : nullptr);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_dcn_bias_grad(const float* dy, float* dbias, int32_t N, int32_t Cout, int32_t npix, float scale, void* stream) {
  RTP_CHECK_ARG(dy && dbias && N > 0 && Cout > 0 && npix > 0, "rtp_dcn_bias_grad: bad arguments");
  dcn_bias_grad_kernel<<<Cout, 256, 0, (cudaStream_t)stream>>>(dy, dbias, N, Cout, npix, scale);
  RTP_LAUNCH_CHECK();
}

// ---- v2 (modulated) ---------------------------------------------------------------------------------------------
extern "C" int rtp_mdcn_fwd(const float* x, const float* offset, const float* mask, const float* w, const float* bias, float* y,
                            int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride,
                            int32_t pad, int32_t dil, int32_t dg, void* stream) {
  RTP_CHECK_ARG(x && offset && mask && w && y, "rtp_mdcn_fwd: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_mdcn_fwd")) return -1;
  return launch_fwd(d, x, offset, mask, w, bias, y, (cudaStream_t)stream);
}

extern "C" int rtp_mdcn_bwd_input(const float* x, const float* offset, const float* mask, const float* w, const float* dy, float* dx,
                                  float* doffset, float* dmask, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout,
                                  int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil, int32_t dg, void* stream) {
  RTP_CHECK_ARG(x && offset && mask && w && dy && dx && doffset && dmask, "rtp_mdcn_bwd_input: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_mdcn_bwd_input")) return -1;
  return launch_bwd_input(d, x, offset, mask, w, dy, dx, doffset, dmask, (cudaStream_t)stream, "rtp_mdcn_bwd_input");
}

extern "C" int rtp_mdcn_bwd_weight(const float* x, const float* offset, const float* mask, const float* dy, float* dw, float* dbias,
                                   int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw,
                                   int32_t stride, int32_t pad, int32_t dil, int32_t dg, float scale, void* stream) {
  RTP_CHECK_ARG(x && offset && mask && dy && dw, "rtp_mdcn_bwd_weight: null pointer");
  Dcn d;
  if (make(d, N, C, H, W, Cout, kh, kw, stride, pad, dil, dg, "rtp_mdcn_bwd_weight")) return -1;
  if (dbias) dcn_bias_grad_kernel<<<Cout, 256, 0, (cudaStream_t)stream>>>(dy, dbias, N, Cout, d.Ho * d.Wo, scale);
  return launch_bwd_weight(d, x, offset, mask, dy, dw, scale, (cudaStream_t)stream);
}
