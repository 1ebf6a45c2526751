// umma_probe.cu — standalone sm_100a probe (no torch): pins the tcgen05 operand-descriptor semantics the
// conv kernels rely on, and measures UMMA issue throughput vs N with SMEM-resident operands.
//
//   T1  K-major SWIZZLE_NONE A with a start address shifted by s*16 B == row shift by s (forward/dgrad taps)
//   T2  MN-major SWIZZLE_NONE A and B (K = positions, MN = channels) with shifted start (wgrad)
//   T3  accumulate flag / TMEM column offsets / N-slices of B by start-address offset (kz stacking)
//   T4  cycles per tcgen05.mma for M=128, N in {32..256}, K-major and MN-major operands
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rtpose_b200/csrc tools/umma_probe.cu -o tools/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc05.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

using namespace tc05;
typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------------
// T1/T3: D[m, n] (+)= sum_{k<K} X[m + shift(step), k] * W[step][n, k];  X: [K/8][P][8], W: [steps][K/8][N][8]
// D written to TMEM columns [col0, col0+N).  Steps alternate shifts from a small table.
__global__ void probe_kmajor(const bf16* __restrict__ X, const bf16* __restrict__ W, float* __restrict__ D, int P,
                             int N, int K, int nsteps, const int* __restrict__ shifts, int col0, int nslice_off,
                             int nslice) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kch = K / 8;
  bf16* sX = reinterpret_cast<bf16*>(smem);
  const uint32_t xbytes = kch * P * 16;
  bf16* sW = reinterpret_cast<bf16*>(smem + ((xbytes + 127) / 128) * 128);
  const uint32_t wbytes_step = kch * N * 16;
  if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_load, xbytes);
    bulk_g2s(sX, X, xbytes, &bar_load);
  }
  for (int i = tid; i < nsteps * kch * N * 8; i += blockDim.x) sW[i] = W[i];
  fence_proxy_async();
  mbar_wait(&bar_load, 0);
  __syncthreads();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = idesc_bf16(128, nslice, 0, 0);
    for (int s = 0; s < nsteps; ++s) {
      for (int k16 = 0; k16 < K / 16; ++k16) {
        uint64_t ad = smem_desc(smem_u32(sX) + shifts[s] * 16 + k16 * 2 * P * 16, /*lbo=*/P * 16, /*sbo=*/128);
        uint64_t bd = smem_desc(smem_u32(sW) + s * wbytes_step + k16 * 2 * N * 16 + nslice_off * 16, /*lbo=*/N * 16,
                                /*sbo=*/128);
        mma_ss(tm + col0, ad, bd, idesc, (s | k16) ? 1u : 0u);
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  for (int c = 0; c < nslice; c += 16) {
    uint32_t v[16];
    tmem_ld16(tm + ((warp * 32u) << 16) + col0 + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * nslice + c + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// ---------------------------------------------------------------------------------------------------
// T2: wgrad-style.  A: Xs [MC = M/8 chunks][P][8] (MN-major, K = positions), B: Ys [NC = N/8][P][8] (MN-major).
// D[m, n] = sum_{k < KP} Xs[m][k + shift] * Ys[n][k], KP positions (multiple of 16).
__global__ void probe_mnmajor(const bf16* __restrict__ X, const bf16* __restrict__ Y, float* __restrict__ D, int P,
                              int N, int KP, int shift) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  bf16* sX = reinterpret_cast<bf16*>(smem);
  bf16* sY = sX + 16 * P * 8;
  if (tid == 0) { mbar_init(&bar_mma, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base);
  for (int i = tid; i < 16 * P * 8; i += blockDim.x) sX[i] = X[i];
  for (int i = tid; i < (N / 8) * P * 8; i += blockDim.x) sY[i] = Y[i];
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N, 1, 1);
    for (int k16 = 0; k16 < KP / 16; ++k16) {
      // MN-major SWIZZLE_NONE: SBO = stride between MN chunks (8 elements), LBO = stride between 8-position K groups
      uint64_t ad = smem_desc(smem_u32(sX) + (shift + k16 * 16) * 16, /*lbo=*/128, /*sbo=*/P * 16);
      uint64_t bd = smem_desc(smem_u32(sY) + (k16 * 16) * 16, /*lbo=*/128, /*sbo=*/P * 16);
      mma_ss(tm, ad, bd, idesc, k16 ? 1u : 0u);
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tm + ((warp * 32u) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * N + c + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// ---------------------------------------------------------------------------------------------------
// T4: throughput.  mode 0: K-major A/B, A start cycles through 9 tap shifts; mode 1: MN-major A/B.
// mode 2: K-major, fixed aligned A.  Each CTA issues `iters` MMAs of shape 128 x N x 16 and reports cycles.
// Interference modes (what a real conv kernel adds around the same MMA stream as mode 0):
//   3: B also cycles through 9 tap slices        4: 3 + a second warp streams bulk copies into shared memory
//   5: 3 + four warps drain TMEM with tcgen05.ld 6: 3 + 4 + 5 together           7: 3 with the D column block rotating
__global__ void probe_rate(long long* __restrict__ cycles, int N, int iters, int mode, const uint8_t* __restrict__ gsrc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_cp[2];
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int P = 600;  // positions per channel chunk
  bf16* sX = reinterpret_cast<bf16*>(smem);                 // [16 chunks][P][8] = 153.6 KB
  bf16* sW = sX + 16 * P * 8;                               // 9 tap slices x [2][N<=96.. 256][8] (mode >= 3: N <= 96)
  uint8_t* sCopy = reinterpret_cast<uint8_t*>(sW) + 9 * 2 * 96 * 16 + 2 * 256 * 16;  // 2 x 16.8 KB landing zone
  for (int i = tid; i < 16 * P * 8 + 9 * 2 * 96 * 8 + 2 * 256 * 8; i += blockDim.x) sX[i] = __float2bfloat16(0.001f * (i % 7));
  if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_cp[0], 1); mbar_init(&bar_cp[1], 1); mbar_fence_init(); stop_flag = 0; }
  if (warp == 0) tmem_alloc<512>(&tmem_base);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  const bool vary_b = mode >= 3, copies = mode == 4 || mode == 6, drain = mode == 5 || mode == 6;
  if (tid == 0) {
    const int Yp = 66;
    const uint32_t idesc = (mode == 1) ? idesc_bf16(128, N, 1, 1) : idesc_bf16(128, N, 0, 0);
    long long t0 = clock64();
    const uint32_t xb = smem_u32(sX), wb = smem_u32(sW);
    for (int it = 0; it < iters; it += 18) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
          uint64_t ad, bd;
          const int sh = (mode == 2) ? 0 : (j / 3) * Yp + (j % 3);
          if (mode == 1) {
            ad = smem_desc(xb + (sh + k16 * 16) * 16, 128, P * 16);
            bd = smem_desc(xb + (k16 * 16 + 200) * 16, 128, P * 16);
          } else {
            ad = smem_desc(xb + sh * 16 + k16 * 2 * P * 16, P * 16, 128);
            bd = vary_b ? smem_desc(wb + (j * 2 * N + k16 * 2 * N) * 16 * 0 + j * 2 * N * 16 + k16 * 0, N * 16, 128) : smem_desc(wb, 256 * 16, 128);
          }
          const uint32_t dcol = (mode == 7) ? tm + ((it / 18) % 4) * 128 : tm;
          mma_ss(dcol, ad, bd, idesc, (it | j | k16) ? 1u : 0u);
        }
      }
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    stop_flag = 1;
  } else if (tid == 32 && copies) {
    // 16.8 KB (4 x 4192 B) per 18 MMAs in the real kernel; here: as fast as the two-slot ring allows
    uint32_t it = 0;
    while (!stop_flag) {
      const int s = it & 1;
      if (it >= 2) mbar_wait(&bar_cp[s], ((it >> 1) - 1) & 1);
      mbar_arrive_expect_tx(&bar_cp[s], 4 * 4192);
      for (int c = 0; c < 4; ++c) bulk_g2s(sCopy + s * 16768 + c * 4192, gsrc + ((size_t)(blockIdx.x * 64 + (it & 63)) * 4 + c) * 4192, 4192, &bar_cp[s]);
      ++it;
    }
    if (it >= 1) mbar_wait(&bar_cp[(it - 1) & 1], ((it - 1) >> 1) & 1);
    if (it >= 2) mbar_wait(&bar_cp[it & 1], ((it - 2) >> 1) & 1);
  } else if (warp >= 2 && drain) {
    // epilogue-like TMEM reads: 96 columns per "plane", back to back
    const uint32_t trow = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    while (!stop_flag) {
      for (int c = 0; c < 96; c += 16) {
        uint32_t v[16];
        tmem_ld16(trow + 256 + c, v);
        tmem_ld_wait();
        sink += v[0];
      }
    }
    if (sink == 0x12345678u) cycles[0] = 0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

// ---------------------------------------------------------------------------------------------------
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float frand() { return (float)(rand() % 2001 - 1000) / 1000.f; }

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s, SMs %d, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  FILE* jf = fopen("gpurun_out/umma_probe.json", "w");
  if (jf) fprintf(jf, "{\"device\": \"%s\", \"sms\": %d", prop.name, prop.multiProcessorCount);
  srand(1);

  // ---------------- T1 / T3
  {
    const int P = 400, K = 32, N = 96, nsteps = 4;
    int h_shifts[nsteps] = {0, 1, 67, 134};
    std::vector<bf16> hX(K / 8 * P * 8), hW(nsteps * (K / 8) * N * 8);
    std::vector<float> fX(P * K), fW(nsteps * N * K);
    for (int p = 0; p < P; ++p) for (int k = 0; k < K; ++k) { float v = bf(frand()); fX[p * K + k] = v; hX[((k / 8) * P + p) * 8 + k % 8] = __float2bfloat16(v); }
    for (int s = 0; s < nsteps; ++s) for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float v = bf(frand()); fW[(s * N + n) * K + k] = v; hW[((s * (K / 8) + k / 8) * N + n) * 8 + k % 8] = __float2bfloat16(v); }
    bf16 *dX, *dW; float* dD; int* dS;
    CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, sizeof(h_shifts)));
    CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dS, h_shifts, sizeof(h_shifts), cudaMemcpyHostToDevice));
    size_t smem = ((K / 8 * P * 16 + 127) / 128) * 128 + nsteps * (K / 8) * N * 16;
    CK(cudaFuncSetAttribute(probe_kmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    struct Case { int col0, noff, nslice; } cases[3] = {{0, 0, 96}, {160, 0, 96}, {64, 32, 64}};
    for (auto c : cases) {
      CK(cudaMemset(dD, 0, 128 * N * 4));
      probe_kmajor<<<1, 128, smem>>>(dX, dW, dD, P, N, K, nsteps, dS, c.col0, c.noff, c.nslice);
      CK(cudaDeviceSynchronize());
      std::vector<float> hD(128 * c.nslice);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < c.nslice; ++n) {
        double ref = 0;
        for (int s = 0; s < nsteps; ++s) for (int k = 0; k < K; ++k) ref += (double)fX[(m + h_shifts[s]) * K + k] * fW[(s * N + n + c.noff) * K + k];
        maxerr = std::max(maxerr, std::fabs(ref - hD[m * c.nslice + n]));
      }
      printf("T1 kmajor shifted-start col0=%d noff=%d nslice=%d: max abs err %.3e %s\n", c.col0, c.noff, c.nslice, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
      if (jf) fprintf(jf, ", \"t1_col%d_n%d_err\": %.3e", c.col0, c.nslice, maxerr);
    }
    cudaFree(dX); cudaFree(dW); cudaFree(dD); cudaFree(dS);
  }
  // ---------------- T2
  {
    const int P = 300, M = 128, KP = 64;
    for (int N : {32, 64}) for (int shift : {0, 1, 67}) {
      std::vector<bf16> hX(M / 8 * P * 8), hY(N / 8 * P * 8);
      std::vector<float> fX(M * P), fY(N * P);
      for (int m = 0; m < M; ++m) for (int p = 0; p < P; ++p) { float v = bf(frand()); fX[m * P + p] = v; hX[((m / 8) * P + p) * 8 + m % 8] = __float2bfloat16(v); }
      for (int n = 0; n < N; ++n) for (int p = 0; p < P; ++p) { float v = bf(frand()); fY[n * P + p] = v; hY[((n / 8) * P + p) * 8 + n % 8] = __float2bfloat16(v); }
      bf16 *dX, *dY; float* dD;
      CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dY, hY.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
      CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY, hY.data(), hY.size() * 2, cudaMemcpyHostToDevice));
      size_t smem = (16 + N / 8) * P * 16;
      CK(cudaFuncSetAttribute(probe_mnmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      probe_mnmajor<<<1, 128, smem>>>(dX, dY, dD, P, N, KP, shift);
      CK(cudaDeviceSynchronize());
      std::vector<float> hD(128 * N);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < KP; ++k) ref += (double)fX[m * P + k + shift] * fY[n * P + k];
        maxerr = std::max(maxerr, std::fabs(ref - hD[m * N + n]));
      }
      printf("T2 mnmajor N=%d shift=%d: max abs err %.3e %s\n", N, shift, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
      if (jf) fprintf(jf, ", \"t2_n%d_s%d_err\": %.3e", N, shift, maxerr);
      cudaFree(dX); cudaFree(dY); cudaFree(dD);
    }
  }
  // ---------------- T4
  {
    const int nsm = prop.multiProcessorCount;
    long long* dC; CK(cudaMalloc(&dC, nsm * 8));
    size_t smem = (16 * 600 * 8 + 9 * 2 * 96 * 8 + 2 * 256 * 8) * 2 + 2 * 16768;
    CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint8_t* gsrc; CK(cudaMalloc(&gsrc, (size_t)nsm * 64 * 4 * 4192 + 4192)); CK(cudaMemset(gsrc, 1, (size_t)nsm * 64 * 4 * 4192));
    const int iters = 18 * 512;
    for (int mode = 0; mode < 8; ++mode) for (int N : {32, 64, 96, 128, 192, 256}) {
      if (mode == 1 && N > 128) continue;
      if (mode >= 3 && N != 96 && N != 32) continue;  // the interference modes model conv_k3s1 (N = 96) and wgrad (N = 32)
      const int threads = mode >= 4 && mode <= 6 ? 192 : 128;
      for (int rep = 0; rep < 2; ++rep) { probe_rate<<<nsm, threads, smem>>>(dC, N, iters, mode, gsrc); CK(cudaDeviceSynchronize()); }
      std::vector<long long> hC(nsm);
      CK(cudaMemcpy(hC.data(), dC, nsm * 8, cudaMemcpyDeviceToHost));
      std::sort(hC.begin(), hC.end());
      double med = (double)hC[nsm / 2] / iters, mx = (double)hC[nsm - 1] / iters;
      double ideal = 128.0 * N / 256.0;  // cycles at 8192 FLOP/clk/SM
      printf("T4 mode=%d N=%3d: %.1f cyc/MMA median (max %.1f), math floor %.0f -> %.0f%% of tensor peak\n", mode, N, med, mx, ideal, 100.0 * ideal / med);
      if (jf) fprintf(jf, ", \"t4_m%d_n%d_cyc\": %.2f", mode, N, med);
    }
    cudaFree(gsrc);
    cudaFree(dC);
  }
  if (jf) { fprintf(jf, "}\n"); fclose(jf); }
  return 0;
}
