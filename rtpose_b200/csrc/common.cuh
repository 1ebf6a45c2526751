// common.cuh — shared device/host helpers for librtpose_b200.so
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/rtpose_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local last error string) -------------------------------------------------
void rtp_set_error(const char* fmt, ...);
#define RTP_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      rtp_set_error(__VA_ARGS__);     \
      return -1;                      \
    }                                 \
  } while (0)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies per device: launchers cache "already opted in up to N bytes"
// per device ordinal so a process that drives several GPUs configures each of them.
#define RTP_MAX_DEVICES 64
static inline int rtp_current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= RTP_MAX_DEVICES) d = 0;
  return d;
}

#define RTP_LAUNCH_CHECK()                                                        \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      rtp_set_error("%s:%d CUDA launch error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                            \
    }                                                                             \
    return 0;                                                                     \
  } while (0)

// ---- P8 geometry ----------------------------------------------------------------------------------------
struct P8 {  // device-side mirror of rtp_p8 with derived pitches
  bf16* ptr;
  int64_t n_stride, c_stride;
  int N, C8, Z, X, Y, Xp, Yp;
  __host__ __device__ P8() {}
  __host__ __device__ explicit P8(const rtp_p8& t)
      : ptr((bf16*)t.ptr), n_stride(t.n_stride), c_stride(t.c_stride), N(t.N), C8(t.C8), Z(t.Z), X(t.X), Y(t.Y),
        Xp(t.X + 2), Yp(t.Y + 2) {}
  // element offset of voxel (z, x, y) channel-chunk 0 inside sample 0 (unpadded coordinates)
  __host__ __device__ int64_t voxel(int z, int x, int y) const {
    return (((int64_t)z * Xp + (x + 1)) * Yp + (y + 1)) * 8;
  }
  __host__ __device__ int64_t plane_elems() const { return (int64_t)Xp * Yp * 8; }
};

// 8 bf16 <-> 8 fp32 through one 16-byte vector
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(bf16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
