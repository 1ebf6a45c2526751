// layout.cu — error plumbing, NCDHW<->P8 conversion at the det3d API boundary, radar-cube ingest,
// weight repacking into UMMA B-operand tiles.
#include <cstdarg>
#include <cstring>
#include <cuda_fp16.h>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
void rtp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* rtp_last_error(void) { return g_err; }
extern "C" int rtp_version(void) { return 100; }
extern "C" int rtp_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

// The L1 / shared-memory split of an SM can only change while the SM is idle, so a kernel that prefers a different split
// than the resident one waits for the SM to drain: the small GroupNorm / reduction kernels (default preference: L1) could
// not start beside the persistent weight-gradient CTAs (196 KB of shared memory), and the main chain stalled behind every
// side-stream weight gradient (bench --timeline: stat_finalize, 4 us of work, started 150 us late).  With one device-wide
// preference every kernel of the process runs under the same (maximum shared memory) split.
int rtp_norm_set_carveout(int pct, int backward_only);  // norm.cu
int rtp_k3s1_set_carveout(int pct);        // conv_k3s1.cu
int rtp_wgrad_k3s1_set_carveout(int pct);  // wgrad_k3s1.cu
extern "C" int rtp_set_shared_carveout(int32_t mode) {
  cudaError_t e = cudaSuccess;
  if (mode == 1) {  // device-wide preference
    e = cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);
  } else if (mode == 2 || mode == 3) {  // 2: the streaming GroupNorm / finalize / reduce kernels; 3: only GroupNorm backward apply + the finalizers
    int r = rtp_norm_set_carveout(cudaSharedmemCarveoutMaxShared, mode == 3);
    if (!r) r = rtp_k3s1_set_carveout(cudaSharedmemCarveoutMaxShared);
    if (!r) r = rtp_wgrad_k3s1_set_carveout(cudaSharedmemCarveoutMaxShared);
    e = (cudaError_t)r;
  } else {
    e = cudaDeviceSetCacheConfig(cudaFuncCachePreferNone);
  }
  if (e != cudaSuccess) {
    rtp_set_error("rtp_set_shared_carveout: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- pack / unpack
// One block: (n, chunk, z, 32-wide x tile, 32-wide y tile); 8 channels x 32 y x 32 x staged in smem so that
// both the NCDHW side (x fastest) and the P8 side (y fastest, 16 B per voxel) are coalesced.
template <bool kPack>
__global__ void __launch_bounds__(256) ncdhw_p8_kernel(float* __restrict__ nc, P8 t, int C, int accumulate) {
  __shared__ float tile[8][32][33];
  const int xt = blockIdx.x * 32, yt = blockIdx.y * 32;
  int b = blockIdx.z;
  const int z = b % t.Z;
  b /= t.Z;
  const int ch = b % t.C8, n = b / t.C8;
  const int tid = threadIdx.x;
  const int64_t vol = (int64_t)t.Z * t.Y * t.X;
  bf16* pbase = t.ptr + n * t.n_stride + ch * t.c_stride;
  if (kPack) {
    for (int i = tid; i < 8 * 32 * 32; i += 256) {
      const int xx = i & 31, yy = (i >> 5) & 31, c = i >> 10;
      const int x = xt + xx, y = yt + yy, cg = ch * 8 + c;
      float v = 0.f;
      if (x < t.X && y < t.Y && cg < C) v = nc[((int64_t)n * C + cg) * vol + ((int64_t)z * t.Y + y) * t.X + x];
      tile[c][yy][xx] = v;
    }
    __syncthreads();
    for (int i = tid; i < 32 * 32; i += 256) {
      const int yy = i & 31, xx = i >> 5;
      const int x = xt + xx, y = yt + yy;
      if (x < t.X && y < t.Y) {
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = tile[c][yy][xx];
        stg16(pbase + t.voxel(z, x, y), pack8(f));
      }
    }
  } else {
    for (int i = tid; i < 32 * 32; i += 256) {
      const int yy = i & 31, xx = i >> 5;
      const int x = xt + xx, y = yt + yy;
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (x < t.X && y < t.Y) unpack8(ldg16(pbase + t.voxel(z, x, y)), f);
#pragma unroll
      for (int c = 0; c < 8; ++c) tile[c][yy][xx] = f[c];
    }
    __syncthreads();
    for (int i = tid; i < 8 * 32 * 32; i += 256) {
      const int xx = i & 31, yy = (i >> 5) & 31, c = i >> 10;
      const int x = xt + xx, y = yt + yy, cg = ch * 8 + c;
      if (x < t.X && y < t.Y && cg < C) {
        float* d = nc + ((int64_t)n * C + cg) * vol + ((int64_t)z * t.Y + y) * t.X + x;
        *d = accumulate ? (*d + tile[c][yy][xx]) : tile[c][yy][xx];
      }
    }
  }
}

static int check_p8(const rtp_p8& t, const char* name) {
  RTP_CHECK_ARG(t.ptr != nullptr, "%s: null pointer", name);
  RTP_CHECK_ARG(t.N > 0 && t.C8 > 0 && t.Z > 0 && t.X > 0 && t.Y > 0, "%s: bad extents", name);
  RTP_CHECK_ARG(((uintptr_t)t.ptr & 15) == 0 && (t.n_stride % 8) == 0 && (t.c_stride % 8) == 0,
                "%s: pointer/strides must be 16-byte aligned", name);
  return 0;
}

extern "C" int rtp_pack_ncdhw(const float* src, rtp_p8 dst, int32_t C, void* stream) {
  if (check_p8(dst, "rtp_pack_ncdhw dst")) return -1;
  RTP_CHECK_ARG(src && C > 0 && C <= dst.C8 * 8, "rtp_pack_ncdhw: bad C=%d for C8=%d", C, dst.C8);
  dim3 grid(ceil_div(dst.X, 32), ceil_div(dst.Y, 32), dst.N * dst.C8 * dst.Z);
  ncdhw_p8_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(src), P8(dst), C, 0);
  RTP_LAUNCH_CHECK();
}
extern "C" int rtp_unpack_ncdhw(rtp_p8 src, float* dst, int32_t C, int32_t accumulate, void* stream) {
  if (check_p8(src, "rtp_unpack_ncdhw src")) return -1;
  RTP_CHECK_ARG(dst && C > 0 && C <= src.C8 * 8, "rtp_unpack_ncdhw: bad C=%d for C8=%d", C, src.C8);
  dim3 grid(ceil_div(src.X, 32), ceil_div(src.Y, 32), src.N * ceil_div(C, 8) * src.Z);
  P8 t(src);
  t.C8 = ceil_div(C, 8);
  ncdhw_p8_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dst, t, C, accumulate);
  RTP_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------- ingest
// raw fp16 [N][D][RZ][RY][RX] -> ROI crop -> (v - a) / s -> clamp -> P8 bf16 [N][D/8][Z][X+2][Y+2][8].
// One block per (n, 8-channel chunk, z, 32x32 (x,y) tile): 8 Doppler planes are read with x fastest
// (64-byte rows, coalesced) and written as 16-byte channel vectors with y fastest.
__global__ void __launch_bounds__(256) ingest_kernel(const __half* __restrict__ raw, int D, int RZ, int RY, int RX,
                                                     int z0, int y0, int x0, float a, float scale, int normalize,
                                                     P8 t, float* __restrict__ f32) {
  __shared__ float tile[8][32][33];
  const int xt = blockIdx.x * 32, yt = blockIdx.y * 32;
  int b = blockIdx.z;
  const int z = b % t.Z;
  b /= t.Z;
  const int ch = b % t.C8, n = b / t.C8;
  const int tid = threadIdx.x;
  for (int i = tid; i < 8 * 32 * 32; i += 256) {
    const int xx = i & 31, yy = (i >> 5) & 31, c = i >> 10;
    const int x = xt + xx, y = yt + yy, d = ch * 8 + c;
    float v = 0.f;
    if (x < t.X && y < t.Y && d < D) {
      v = __half2float(raw[((((int64_t)n * D + d) * RZ + (z0 + z)) * RY + (y0 + y)) * RX + (x0 + x)]);
      if (normalize) {
        // reference arithmetic: (float32(v) - start) / scale ; then clamp (cruw_pose.py:182-183)
        v = (v - a) / scale;  // IEEE division: bit-identical to numpy's float32 arithmetic
        v = v < 0.f ? 0.f : v;
      }
      if (f32) f32[((((int64_t)n * D + d) * t.Z + z) * t.Y + y) * t.X + x] = v;
    }
    tile[c][yy][xx] = v;
  }
  __syncthreads();
  bf16* pbase = t.ptr + n * t.n_stride + ch * t.c_stride;
  for (int i = tid; i < 32 * 32; i += 256) {
    const int yy = i & 31, xx = i >> 5;
    const int x = xt + xx, y = yt + yy;
    if (x < t.X && y < t.Y) {
      float f[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = tile[c][yy][xx];
      stg16(pbase + t.voxel(z, x, y), pack8(f));
    }
  }
}

// Vectorised variant (no fp32 side output, 16-byte aligned raw rows): one block per (n, 8-Doppler chunk, z, 8 y rows) and the
// WHOLE x range of the ROI.  Reads are 16-byte vectors of 8 halves from the aligned superset [x0 & ~7, ...) of each raw row
// (one coalesced run of ~340 B per (plane, row) instead of 2-byte scalar loads of 64-byte pieces: the scalar kernel ran at
// 0.21 of the HBM rate, instruction-bound), normalised / clamped / rounded to bf16 into shared memory, then written as
// 16-byte channel vectors, 8 consecutive y (one full 128-byte line) per x.  Same arithmetic, same rounding.
constexpr int kIngY = 8, kIngXMax = 176;  // rows per block; widest ROI handled (vectors cover <= kIngXMax + 8 halves)
__global__ void __launch_bounds__(256) ingest_vec_kernel(const __half* __restrict__ raw, int D, int RZ, int RY, int RX,
                                                         int z0, int y0, int x0, float a, float scale, int normalize, P8 t) {
  // row pitch 200 halves = 400 B: 16-byte aligned for the vector stores, and 100 words % 32 = 4, so the 8 lanes that read the
  // same x of 8 consecutive rows in the write phase hit 8 different banks (a 192-half pitch put them all in one)
  __shared__ __align__(16) bf16 tile[8][kIngY][kIngXMax + 24];
  const int yt = blockIdx.x * kIngY;
  int b = blockIdx.y;
  const int z = b % t.Z;
  b /= t.Z;
  const int ch = b % t.C8, n = b / t.C8;
  const int xa = x0 & ~7;                          // aligned start of the superset
  const int nvec = (x0 + t.X - xa + 7) >> 3;       // 16-byte vectors per row
  const int tid = threadIdx.x;
  for (int i = tid; i < 8 * kIngY * nvec; i += 256) {
    const int v = i % nvec, r = i / nvec, yy = r % kIngY, c = r / kIngY;
    const int y = yt + yy, d = ch * 8 + c;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    const bool ok = y < t.Y && d < D;
    if (ok) q = __ldg(reinterpret_cast<const uint4*>(raw + ((((int64_t)n * D + d) * RZ + (z0 + z)) * RY + (y0 + y)) * RX + xa) + v);
    const __half2* h2 = reinterpret_cast<const __half2*>(&q);
    __nv_bfloat162 o2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __half22float2(h2[k]);
      if (normalize) {
        f.x = (f.x - a) / scale; f.x = f.x < 0.f ? 0.f : f.x;
        f.y = (f.y - a) / scale; f.y = f.y < 0.f ? 0.f : f.y;
      }
      if (!ok) f.x = f.y = 0.f;
      o2[k] = __floats2bfloat162_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(&tile[c][yy][v * 8]) = *reinterpret_cast<const uint4*>(o2);
  }
  __syncthreads();
  bf16* pbase = t.ptr + n * t.n_stride + ch * t.c_stride;
  const int xs = x0 - xa;  // first ROI element inside the superset
  for (int i = tid; i < t.X * kIngY; i += 256) {
    const int yy = i % kIngY, x = i / kIngY, y = yt + yy;
    if (y < t.Y) {
      bf16x8 o;
      bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
      for (int c = 0; c < 8; ++c) ob[c] = tile[c][yy][xs + x];
      stg16(pbase + t.voxel(z, x, y), *reinterpret_cast<const uint4*>(&o));
    }
  }
}

extern "C" int rtp_ingest_pack(const void* raw_f16, int32_t N, int32_t D, int32_t RZ, int32_t RY, int32_t RX,
                               int32_t z0, int32_t y0, int32_t x0, float norm_start, float norm_scale,
                               int32_t normalize, rtp_p8 dst, float* dst_f32, void* stream) {
  if (check_p8(dst, "rtp_ingest_pack dst")) return -1;
  RTP_CHECK_ARG(raw_f16 && N == dst.N && D <= dst.C8 * 8, "rtp_ingest_pack: N/D mismatch");
  RTP_CHECK_ARG(z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + dst.Z <= RZ && y0 + dst.Y <= RY && x0 + dst.X <= RX,
                "rtp_ingest_pack: ROI outside the raw cube");
  RTP_CHECK_ARG(!normalize || norm_scale != 0.f, "rtp_ingest_pack: zero normalisation scale");
  const int64_t blocks_y = (int64_t)dst.N * dst.C8 * dst.Z;
  if (!dst_f32 && RX % 8 == 0 && ((uintptr_t)raw_f16 & 15) == 0 && (x0 & 7) + dst.X <= kIngXMax + 8 &&
      (x0 & ~7) + (((x0 & 7) + dst.X + 7) & ~7) <= RX && blocks_y <= 65535) {
    ingest_vec_kernel<<<dim3(ceil_div(dst.Y, kIngY), (unsigned)blocks_y), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)raw_f16, D, RZ, RY, RX, z0, y0, x0, norm_start, normalize ? norm_scale : 1.f, normalize, P8(dst));
    RTP_LAUNCH_CHECK();
  }
  dim3 grid(ceil_div(dst.X, 32), ceil_div(dst.Y, 32), dst.N * dst.C8 * dst.Z);
  ingest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)raw_f16, D, RZ, RY, RX, z0, y0, x0, norm_start,
                                                        normalize ? norm_scale : 1.f, normalize, P8(dst),
                                                        dst_f32);
  RTP_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------- weights
// w fp32 [Cout][Cin][ntaps] -> dst bf16 [tap][KP/8][NP][8]
//   mode 0: k = ci - ci0 (ci in [ci0, ci0+ci_n)), n = co ;  mode 1: k = co, n = ci - ci0
__device__ __forceinline__ void weight_pack_elems(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int ntaps,
                                                  int ci0, int ci_n, int KP, int NP, int mode, int64_t first, int64_t stride) {
  const int64_t total = (int64_t)ntaps * KP * NP;
  for (int64_t i = first; i < total; i += stride) {
    const int k8 = i & 7;
    int64_t r = i >> 3;
    const int n = r % NP;
    r /= NP;
    const int kc = r % (KP / 8);
    const int tap = r / (KP / 8);
    const int k = kc * 8 + k8;
    int co, ci;
    if (mode == 0) { ci = k; co = n; } else { co = k; ci = n; }
    float v = 0.f;
    if (co < Cout && ci < ci_n) v = w[((int64_t)co * Cin + (ci0 + ci)) * ntaps + tap];
    dst[i] = __float2bfloat16(v);
  }
}
__global__ void weight_pack_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int ntaps,
                                   int ci0, int ci_n, int KP, int NP, int mode) {
  weight_pack_elems(w, dst, Cout, Cin, ntaps, ci0, ci_n, KP, NP, mode, blockIdx.x * (int64_t)blockDim.x + threadIdx.x,
                    (int64_t)gridDim.x * blockDim.x);
}
extern "C" int rtp_weight_pack(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin, int32_t ntaps, int32_t ci0,
                               int32_t ci_n, int32_t KP, int32_t NP, int32_t mode, void* stream) {
  RTP_CHECK_ARG(w && dst_bf16, "rtp_weight_pack: null pointer");
  RTP_CHECK_ARG(KP % 16 == 0 && NP % 16 == 0 && NP <= 256, "rtp_weight_pack: KP/NP must be multiples of 16");
  RTP_CHECK_ARG(ci0 >= 0 && ci_n > 0 && ci0 + ci_n <= Cin, "rtp_weight_pack: bad input-channel slice");
  RTP_CHECK_ARG(mode == 0 ? (ci_n <= KP && Cout <= NP) : (Cout <= KP && ci_n <= NP), "rtp_weight_pack: padding too small");
  const int64_t total = (int64_t)ntaps * KP * NP;
  weight_pack_kernel<<<ceil_div(total, 256) > 1024 ? 1024 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w, (bf16*)dst_bf16, Cout, Cin, ntaps, ci0, ci_n, KP, NP, mode);
  RTP_LAUNCH_CHECK();
}

// k3s1 plane-streaming pack: dst[(ky,kx) tap 9][KP/8][3*NPo][8]; N index = j*NPo + co with j = 2 - kz
// (block order [kz=2 | kz=1 | kz=0] so the three TMEM accumulator blocks of output planes z-1, z, z+1 are
// contiguous).  In-plane tap index t9 = kx*3 + ky (x is the slow in-plane axis of the P8 layout).
// transpose_flip = 1 builds the dgrad operand: K = Cout, N = Cin, taps mirrored.
// (Cin_total, ci0): the pack may cover a window [ci0, ci0 + Cin) of the weight's input channels (a group of the
// space-to-depth dgrad) without the caller materialising the slice.
__device__ __forceinline__ void weight_pack_k3s1_elems(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int KP,
                                                       int NPo, int tf, int Cin_total, int ci0, int64_t first, int64_t stride) {
  const int N3 = 3 * NPo;
  const int64_t total = (int64_t)9 * KP * N3;
  for (int64_t i = first; i < total; i += stride) {
    const int k8 = i & 7;
    int64_t r = i >> 3;
    const int n = r % N3;
    r /= N3;
    const int kc = r % (KP / 8);
    const int t9 = r / (KP / 8);
    const int k = kc * 8 + k8;
    const int j = n / NPo, nn = n % NPo;
    int kz = 2 - j, kx = t9 / 3, ky = t9 % 3;
    int co, ci;
    if (!tf) { ci = k; co = nn; } else { co = k; ci = nn; kz = 2 - kz; ky = 2 - ky; kx = 2 - kx; }
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((int64_t)co * Cin_total + ci0 + ci) * 27 + (kz * 3 + ky) * 3 + kx];
    dst[i] = __float2bfloat16(v);
  }
}
__global__ void weight_pack_k3s1_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int KP,
                                        int NPo, int tf, int Cin_total, int ci0) {
  weight_pack_k3s1_elems(w, dst, Cout, Cin, KP, NPo, tf, Cin_total, ci0, blockIdx.x * (int64_t)blockDim.x + threadIdx.x,
                         (int64_t)gridDim.x * blockDim.x);
}
// Every pack of a training step in ONE launch: block b serves the job whose block range [block0, block0 + nblocks) holds b
// (binary search over the table).  ~130 separate 2-us launches formed a 0.34 ms serial chain in front of the first conv.
__global__ void __launch_bounds__(256) weight_pack_batch_kernel(const rtp_pack_job* __restrict__ jobs, int njobs) {
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= b) lo = mid; else hi = mid - 1;
  }
  const rtp_pack_job j = jobs[lo];
  const int64_t first = (int64_t)(b - j.block0) * 256 + threadIdx.x, stride = (int64_t)j.nblocks * 256;
  if (j.kind == 0)
    weight_pack_elems((const float*)j.w, (bf16*)j.dst, j.Cout, j.Cin_total, j.ntaps, j.ci0, j.ci_n, j.KP, j.NP, j.flag, first, stride);
  else
    weight_pack_k3s1_elems((const float*)j.w, (bf16*)j.dst, j.Cout, j.ci_n, j.KP, j.NP, j.flag, j.Cin_total, j.ci0, first, stride);
}
static int weight_pack_k3s1_launch(const float* w, void* dst_bf16, int Cout, int Cin, int KP, int NPo, int transpose_flip,
                                   int Cin_total, int ci0, void* stream) {
  RTP_CHECK_ARG(w && dst_bf16, "rtp_weight_pack_k3s1: null pointer");
  RTP_CHECK_ARG(KP % 16 == 0 && NPo % 16 == 0 && 3 * NPo <= 256, "rtp_weight_pack_k3s1: bad KP/NPo");
  RTP_CHECK_ARG(transpose_flip ? (Cout <= KP && Cin <= NPo) : (Cin <= KP && Cout <= NPo),
                "rtp_weight_pack_k3s1: padding too small");
  RTP_CHECK_ARG(ci0 >= 0 && Cin >= 1 && ci0 + Cin <= Cin_total, "rtp_weight_pack_k3s1: bad input-channel window");
  const int64_t total = (int64_t)9 * KP * 3 * NPo;
  weight_pack_k3s1_kernel<<<ceil_div(total, 256) > 1024 ? 1024 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w, (bf16*)dst_bf16, Cout, Cin, KP, NPo, transpose_flip, Cin_total, ci0);
  RTP_LAUNCH_CHECK();
}
extern "C" int rtp_weight_pack_batch(const rtp_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream) {
  RTP_CHECK_ARG(jobs_dev && njobs > 0 && total_blocks > 0, "rtp_weight_pack_batch: bad arguments");
  weight_pack_batch_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_dev, njobs);
  RTP_LAUNCH_CHECK();
}
extern "C" int rtp_weight_pack_k3s1(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin, int32_t KP, int32_t NPo,
                                    int32_t transpose_flip, void* stream) {
  return weight_pack_k3s1_launch(w, dst_bf16, Cout, Cin, KP, NPo, transpose_flip, Cin, 0, stream);
}
extern "C" int rtp_weight_pack_k3s1_window(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin_total, int32_t ci0,
                                           int32_t ci_n, int32_t KP, int32_t NPo, int32_t transpose_flip, void* stream) {
  return weight_pack_k3s1_launch(w, dst_bf16, Cout, ci_n, KP, NPo, transpose_flip, Cin_total, ci0, stream);
}

// ---------------------------------------------------------------------------------------------- space-to-depth weights
// A stride-2 3x3x3 conv (pad 1) over x equals a stride-1 3x3x3 conv over the space-to-depth view of x (8 parity groups
// of Cin channels at half resolution, see norm.cu s2d_offset): input index 2o + k - 1 has parity 1 / offset -1 for
// k = 0, parity 0 / offset 0 for k = 1, parity 1 / offset 0 for k = 2.  W'[co][par*Cin + ci][t] with taps t in the
// reference order (kz*3 + ky)*3 + kx holds W[co][ci][k] at the 27 matching (parity, offset) pairs and zero elsewhere.
__device__ __forceinline__ int s2d_src_tap(int par, int t) { return par == 0 ? (t == 1 ? 1 : -1) : (t == 0 ? 0 : (t == 1 ? 2 : -1)); }

__global__ void weight_s2d_expand_kernel(const float* __restrict__ w, float* __restrict__ we, int Cout, int Cin) {
  const int64_t total = (int64_t)Cout * 8 * Cin * 27;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % 27);
    const int64_t r = i / 27;
    const int kk = (int)(r % (8 * Cin)), co = (int)(r / (8 * Cin));
    const int par = kk / Cin, ci = kk % Cin;
    const int sz = s2d_src_tap((par >> 2) & 1, t / 9), sy = s2d_src_tap(par & 1, (t / 3) % 3), sx = s2d_src_tap((par >> 1) & 1, t % 3);
    we[i] = (sz >= 0 && sy >= 0 && sx >= 0) ? w[((int64_t)co * Cin + ci) * 27 + (sz * 3 + sy) * 3 + sx] : 0.f;
  }
}
// the transpose: dW[co][ci][k] (=|+=) dW'[co][par(k)*Cin + ci][t(k)]
__global__ void weight_s2d_fold_kernel(const float* __restrict__ dwe, float* __restrict__ dw, int Cout, int Cin, int accumulate) {
  const int64_t total = (int64_t)Cout * Cin * 27;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % 27);
    const int64_t r = i / 27;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
    // k = 0 -> (parity 1, tap 0); 1 -> (0, 1); 2 -> (1, 1)
    const int pz = kz != 1, py = ky != 1, px = kx != 1;
    const int tz = kz == 0 ? 0 : 1, ty = ky == 0 ? 0 : 1, tx = kx == 0 ? 0 : 1;
    const int par = (pz << 2) | (px << 1) | py;
    const float v = dwe[((int64_t)co * 8 * Cin + par * Cin + ci) * 27 + (tz * 3 + ty) * 3 + tx];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}
extern "C" int rtp_weight_s2d_expand(const float* w, float* w_s2d, int32_t Cout, int32_t Cin, void* stream) {
  RTP_CHECK_ARG(w && w_s2d && Cout > 0 && Cin > 0, "rtp_weight_s2d_expand: bad arguments");
  const int64_t total = (int64_t)Cout * 8 * Cin * 27;
  weight_s2d_expand_kernel<<<ceil_div(total, 256) > 1024 ? 1024 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, w_s2d, Cout, Cin);
  RTP_LAUNCH_CHECK();
}
extern "C" int rtp_weight_s2d_fold(const float* dw_s2d, float* dw, int32_t Cout, int32_t Cin, int32_t accumulate, void* stream) {
  RTP_CHECK_ARG(dw_s2d && dw && Cout > 0 && Cin > 0, "rtp_weight_s2d_fold: bad arguments");
  const int64_t total = (int64_t)Cout * Cin * 27;
  weight_s2d_fold_kernel<<<ceil_div(total, 256) > 1024 ? 1024 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dw_s2d, dw, Cout, Cin,
                                                                                                                     accumulate);
  RTP_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------- misc
__global__ void scale_kernel(float* __restrict__ b, int64_t n, float s) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] *= s;
}
extern "C" int rtp_scale_f32(float* buf, int64_t n, float scale, void* stream) {
  RTP_CHECK_ARG(buf && n >= 0, "rtp_scale_f32: bad args");
  if (n == 0) return 0;
  scale_kernel<<<ceil_div(n, 1024) > 592 ? 592 : ceil_div(n, 1024), 256, 0, (cudaStream_t)stream>>>(buf, n, scale);
  RTP_LAUNCH_CHECK();
}
