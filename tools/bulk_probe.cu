// bulk_probe.cu — per-SM throughput of cp.async.bulk global->shared as a function of copy size, issue pattern and
// pipeline depth.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_probe tools/bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rtpose_b200/csrc/tc05.cuh"
using namespace tc05;

struct Args { const uint8_t* src; size_t span; int copy_bytes, ncopy, stages, iters, parallel; size_t stride; };

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_fence_init(); }
  __syncthreads();
  const uint32_t step_bytes = (uint32_t)a.copy_bytes * a.ncopy;
  if (warp == 0) {
    for (int it = 0; it < a.iters; ++it) {
      const int s = it % a.stages;
      if (lane == 0) { mbar_wait(&empty[s], ((it / a.stages) & 1) ^ 1); mbar_arrive_expect_tx(&full[s], step_bytes); }
      __syncwarp();
      const size_t base = ((size_t)(blockIdx.x * a.iters + it) * a.ncopy);
      if (a.parallel) {
        for (int i = lane; i < a.ncopy; i += 32)
          bulk_g2s(smem + (size_t)s * step_bytes + (size_t)i * a.copy_bytes, a.src + ((base + i) * a.stride) % a.span, a.copy_bytes, &full[s]);
      } else if (lane == 0) {
        for (int i = 0; i < a.ncopy; ++i)
          bulk_g2s(smem + (size_t)s * step_bytes + (size_t)i * a.copy_bytes, a.src + ((base + i) * a.stride) % a.span, a.copy_bytes, &full[s]);
      }
    }
  } else if (lane == 0) {
    for (int it = 0; it < a.iters; ++it) {
      const int s = it % a.stages;
      mbar_wait(&full[s], (it / a.stages) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

int main() {
  const size_t big = (size_t)2 << 30;
  uint8_t* buf; cudaMalloc(&buf, big + (1 << 20)); cudaMemset(buf, 1, big);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%-8s %-6s %-6s %-7s %-9s %-5s %10s %10s\n", "copy_B", "ncopy", "stages", "issue", "footprint", "iters", "TB/s chip", "B/clk/SM");
  const int sizes[] = {1024, 2048, 4096, 8192, 16384};
  for (int foot = 0; foot < 2; ++foot)
    for (int par = 0; par < 2; ++par)
      for (int stages : {2, 4})
        for (int cb : sizes) {
          Args a; a.src = buf; a.span = foot ? big : ((size_t)48 << 20); a.copy_bytes = cb; a.ncopy = 32768 / cb; a.stages = stages;
          a.iters = 400; a.parallel = par; a.stride = cb + 4096 * 3;  // scattered pieces like channel-chunk rows
          a.stride = (a.stride + 15) / 16 * 16;
          const size_t smem = (size_t)stages * 32768;
          probe<<<148, 64, smem>>>(a); cudaDeviceSynchronize();
          cudaEventRecord(e0);
          probe<<<148, 64, smem>>>(a);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = 148.0 * a.iters * 32768.0;
          int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
          printf("%-8d %-6d %-6d %-7s %-9s %-5d %10.2f %10.1f\n", cb, a.ncopy, stages, par ? "warp" : "lane0", foot ? "2GB(HBM)" : "48MB(L2)", a.iters,
                 bytes / ms / 1e9, bytes / 148 / (ms * 1e-3 * clk * 1e3));
        }
  cudaError_t e = cudaGetLastError(); if (e) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
