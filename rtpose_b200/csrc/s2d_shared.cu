// s2d_shared.cu — sibling stride-2 exchange convs share ONE normalised space-to-depth view.
//
// In an HR module the full-resolution branch output x0 feeds up to three `GroupNorm -> conv3x3x3 stride 2` fuse layers
// (hr_util/hr3d.py:135-200, fuse_layers[i][0][0], i = 1..3).  Their GroupNorms share the statistics of x0 and differ only in
// the affine (gamma_k, beta_k), so  xn_k = gamma_k * xhat + beta_k  with ONE xhat = (x0 - mean) * rstd.  Each of them used to
// write its own normalised view (336 MB of traffic), and its backward re-read x0 twice (reduction + apply) and
// read-modify-wrote dL/dx0.  With the affine folded into the conv,
//
//     conv_W(gamma*xhat + beta*1_inside) = conv_{W*diag(gamma)}(xhat) + B[co][class(pos)],
//     B[co][cls] = sum_ci beta_ci * sum_{taps inside the volume for border class cls} W[co][ci][tap],
//
// the siblings read the same view of xhat, their dgrads (with the gamma-folded weights) ACCUMULATE into one dL/dxhat view, and
// one GroupNorm backward (gamma = 1) finishes all of them.  The class of an output position is which of its coordinates are 0:
// input index 2o + k - 1 leaves the volume only for o = 0, k = 0 (extents are even), so there are 8 classes.  The parameter
// gradients come from the weight gradient over xhat, dW' = wgrad(xhat, dy):
//
//     dW[co][ci][t]  = gamma_ci * dW'[co][ci][t] + beta_ci * T[co][t],     T[co][t] = sum_{pos : tap t inside} dy[co][pos]
//     dgamma[ci]     = sum_{co,t} W[co][ci][t] * dW'[co][ci][t]
//     dbeta[ci]      = sum_{co,t} W[co][ci][t] * T[co][t]
//
// where T follows by inclusion-exclusion from the 8 "coordinate == 0" box sums S_A of dy (rtp_s2d_box_sums).  Everything is
// summed in a fixed order (run-to-run identical).
#include "common.cuh"

namespace {

__device__ __forceinline__ int src_tap(int par, int t) { return par == 0 ? (t == 1 ? 1 : -1) : (t == 0 ? 0 : (t == 1 ? 2 : -1)); }

// we[co][par*Cin + ci][t] = W[co][ci][k(par, t)] * gamma[ci]  (zero where (par, t) matches no tap)
__global__ void fold_expand_kernel(const float* __restrict__ w, const float* __restrict__ gamma, float* __restrict__ we, int Cout,
                                   int Cin) {
  const int64_t total = (int64_t)Cout * 8 * Cin * 27;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % 27);
    const int64_t r = i / 27;
    const int kk = (int)(r % (8 * Cin)), co = (int)(r / (8 * Cin));
    const int par = kk / Cin, ci = kk % Cin;
    const int sz = src_tap((par >> 2) & 1, t / 9), sy = src_tap(par & 1, (t / 3) % 3), sx = src_tap((par >> 1) & 1, t % 3);
    we[i] = (sz >= 0 && sy >= 0 && sx >= 0) ? w[((int64_t)co * Cin + ci) * 27 + (sz * 3 + sy) * 3 + sx] * gamma[ci] : 0.f;
  }
}

// bias_cls[cls][co] (cls bit 2: zo == 0, bit 1: xo == 0, bit 0: yo == 0): one block per co, fixed-order tree
__global__ void __launch_bounds__(256) bias_cls_kernel(const float* __restrict__ w, const float* __restrict__ beta,
                                                       float* __restrict__ bias_cls, int Cin, int Cout) {
  __shared__ float sh[256][8];
  const int co = blockIdx.x;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int i = threadIdx.x; i < Cin * 27; i += 256) {
    const int ci = i / 27, k = i % 27;
    const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
    const float v = beta[ci] * w[((int64_t)co * Cin + ci) * 27 + k];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool outside = ((c & 4) && kz == 0) || ((c & 2) && kx == 0) || ((c & 1) && ky == 0);
      if (!outside) acc[c] += v;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) sh[threadIdx.x][c] = acc[c];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int c = 0; c < 8; ++c) sh[threadIdx.x][c] += sh[threadIdx.x + s][c];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) bias_cls[threadIdx.x * Cout + co] = sh[0][threadIdx.x];
}

// r[0][c][z][x][y] = bf16(bias_cls[cls(z, x, y)][c] - bias_cls[0][c]): zero except on the three low faces
__global__ void __launch_bounds__(256) border_bias_kernel(const float* __restrict__ bias_cls, P8 r, int C) {
  const int64_t total = (int64_t)r.C8 * r.Z * r.X * r.Y;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    uint32_t q = (uint32_t)i;
    const int y = (int)(q % (uint32_t)r.Y); q /= (uint32_t)r.Y;
    const int x = (int)(q % (uint32_t)r.X); q /= (uint32_t)r.X;
    const int z = (int)(q % (uint32_t)r.Z);
    const int c8 = (int)(q / (uint32_t)r.Z);
    const int cls = (z == 0 ? 4 : 0) | (x == 0 ? 2 : 0) | (y == 0 ? 1 : 0);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      f[j] = (c < C && cls) ? bias_cls[cls * C + c] - bias_cls[c] : 0.f;
    }
    stg16(r.ptr + c8 * r.c_stride + r.voxel(z, x, y), pack8(f));
  }
}

constexpr int kBoxSlabs = 16;

// partial[n][c8][slab][A][8]: sums of dy over the voxels whose coordinates are 0 on every axis in A (A bit 2: z, 1: x, 0: y)
__global__ void __launch_bounds__(256) box_sums_partial_kernel(P8 dy, float* __restrict__ partial) {
  __shared__ float sh[8][64];
  const int slab = blockIdx.x, c8 = blockIdx.y, n = blockIdx.z;
  const int R = dy.Z * dy.X;
  const int r0 = (int)((int64_t)R * slab / kBoxSlabs), r1 = (int)((int64_t)R * (slab + 1) / kBoxSlabs);
  int log2ty = 3;
  while ((1 << log2ty) < dy.Y && log2ty < 6) ++log2ty;
  const int TY = 1 << log2ty, ty = threadIdx.x & (TY - 1), tr = threadIdx.x >> log2ty, rstep = 256 >> log2ty;
  const bf16* base = dy.ptr + n * dy.n_stride + c8 * dy.c_stride;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
  for (int row = r0 + tr; row < r1; row += rstep) {
    const int z = row / dy.X, x = row - z * dy.X;
    for (int y = ty; y < dy.Y; y += TY) {
      float f[8];
      unpack8(ldg16(base + dy.voxel(z, x, y)), f);
      const int zero = (z == 0 ? 4 : 0) | (x == 0 ? 2 : 0) | (y == 0 ? 1 : 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
      if (zero) {  // voxels on a low face only (64 predicated adds per vector for everyone cost 81 us per launch)
#pragma unroll
        for (int a = 1; a < 8; ++a) {
          if ((a & zero) == a) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][j] += f[j];
          }
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = warp_sum(acc[a][j]);
      if (lane == 0) sh[warp][a * 8 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    partial[(((size_t)n * dy.C8 + c8) * kBoxSlabs + slab) * 64 + threadIdx.x] = s;
  }
}

// One block per output channel co: S[co][A] = fixed-order sum of the box-sum partials (warp A, fp64), T[co][tap] by
// inclusion-exclusion (also written to Tg for the gamma / beta kernel), then dW[co][:][:] (see the file header).  (A single
// 256-thread block did all of this in 116 us per launch, on the weight-gradient stream of every shared stride-2 conv.)
__global__ void __launch_bounds__(256) fold_wgrad_dw_kernel(const float* __restrict__ dwp, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ partial, int N,
                                                            int C8dy, float* __restrict__ dw, float* __restrict__ Tg, int Cin,
                                                            int acc_w) {
  __shared__ float S[8], T[27];
  const int co = blockIdx.x, c8 = co >> 3, j = co & 7;
  const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    double s = 0;
    const int total = N * kBoxSlabs;  // (n, slab) pairs, lane-strided, then a fixed xor tree
    for (int i = lane; i < total; i += 32) {
      const int n = i / kBoxSlabs, slab = i - n * kBoxSlabs;
      s += partial[(((size_t)n * C8dy + c8) * kBoxSlabs + slab) * 64 + a * 8 + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) S[a] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    const int k = threadIdx.x;
    const int kz = k / 9, ky = (k / 3) % 3, kx = k % 3;
    const int m = (kz == 0 ? 4 : 0) | (kx == 0 ? 2 : 0) | (ky == 0 ? 1 : 0);  // axes on which the tap needs o >= 1
    float t = 0.f;
    for (int b = 0; b < 8; ++b) {
      if ((b & m) != b) continue;
      const float v = S[b];
      t += (__popc(b) & 1) ? -v : v;
    }
    T[k] = t;
    Tg[co * 27 + k] = t;
  }
  __syncthreads();
  const int total = Cin * 27;
  const size_t base = (size_t)co * total;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int k = i % 27, ci = i / 27;
    const float v = gamma[ci] * dwp[base + i] + beta[ci] * T[k];
    dw[base + i] = acc_w ? dw[base + i] + v : v;
  }
}

// dgamma[ci] / dbeta[ci]: one warp per input channel, lane-strided over the Cout * 27 products (fp64), fixed xor tree
__global__ void __launch_bounds__(256) fold_wgrad_gb_kernel(const float* __restrict__ dwp, const float* __restrict__ w,
                                                            const float* __restrict__ Tg, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int Cout, int Cin, int acc_gb) {
  const int lane = threadIdx.x & 31;
  const int ci = blockIdx.x * 8 + ((int)threadIdx.x >> 5);
  if (ci >= Cin) return;
  double g = 0, b = 0;
  for (int i = lane; i < Cout * 27; i += 32) {
    const int co = i / 27, k = i - co * 27;
    const float wv = w[((int64_t)co * Cin + ci) * 27 + k];
    g += (double)wv * (double)dwp[((int64_t)co * Cin + ci) * 27 + k];
    b += (double)wv * (double)Tg[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    dgamma[ci] = acc_gb ? dgamma[ci] + (float)g : (float)g;
    dbeta[ci] = acc_gb ? dbeta[ci] + (float)b : (float)b;
  }
}

}  // namespace

extern "C" int rtp_s2d_fold_weights(const float* w, const float* gamma, const float* beta, float* w_s2d, float* bias_cls,
                                    int32_t Cout, int32_t Cin, void* stream) {
  RTP_CHECK_ARG(w && gamma && beta && w_s2d && bias_cls && Cout > 0 && Cin > 0, "rtp_s2d_fold_weights: bad arguments");
  const int64_t total = (int64_t)Cout * 8 * Cin * 27;
  fold_expand_kernel<<<ceil_div(total, 256) > 1024 ? 1024 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, gamma, w_s2d,
                                                                                                               Cout, Cin);
  bias_cls_kernel<<<Cout, 256, 0, (cudaStream_t)stream>>>(w, beta, bias_cls, Cin, Cout);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_s2d_border_bias(const float* bias_cls, rtp_p8 r, int32_t C, void* stream) {
  RTP_CHECK_ARG(bias_cls && r.ptr && C > 0 && C <= r.C8 * 8, "rtp_s2d_border_bias: bad arguments");
  P8 t(r);
  t.C8 = ceil_div(C, 8);
  const int64_t total = (int64_t)t.C8 * t.Z * t.X * t.Y;
  border_bias_kernel<<<ceil_div(total, 256) > 1184 ? 1184 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(bias_cls, t, C);
  RTP_LAUNCH_CHECK();
}

// box-sum partials [N][C8][kBoxSlabs][64], then T[Cout <= 256][27]
extern "C" int64_t rtp_s2d_box_sums_workspace_bytes(int32_t N, int32_t C8) { return ((int64_t)N * C8 * kBoxSlabs * 64 + 256 * 27) * 4; }

extern "C" int rtp_s2d_fold_wgrad(rtp_p8 dy, const float* dw_xhat, const float* w, const float* gamma, const float* beta,
                                  float* workspace, float* dW, float* dgamma, float* dbeta, int32_t Cout, int32_t Cin,
                                  int32_t accumulate_w, int32_t accumulate_gb, void* stream) {
  RTP_CHECK_ARG(dy.ptr && dw_xhat && w && gamma && beta && workspace && dW && dgamma && dbeta, "rtp_s2d_fold_wgrad: null argument");
  RTP_CHECK_ARG(Cout > 0 && Cout <= dy.C8 * 8 && Cin > 0 && Cout <= 256, "rtp_s2d_fold_wgrad: bad channel counts");
  P8 t(dy);
  t.C8 = ceil_div(Cout, 8);
  box_sums_partial_kernel<<<dim3(kBoxSlabs, t.C8, t.N), 256, 0, (cudaStream_t)stream>>>(t, workspace);
  float* Tg = workspace + (size_t)t.N * t.C8 * kBoxSlabs * 64;
  fold_wgrad_dw_kernel<<<Cout, 256, 0, (cudaStream_t)stream>>>(dw_xhat, gamma, beta, workspace, t.N, t.C8, dW, Tg, Cin, accumulate_w);
  fold_wgrad_gb_kernel<<<ceil_div(Cin, 8), 256, 0, (cudaStream_t)stream>>>(dw_xhat, w, Tg, dgamma, dbeta, Cout, Cin, accumulate_gb);
  RTP_LAUNCH_CHECK();
}
