// train_aux.cu — the two "next" rows of SURVEY.md §8f that sit directly on either side of the hot path:
//   N1  CenterNet target assignment on the device (gaussian heat-map splat + ind/mask/cat/anno_pose), replacing the
//       numpy code in the DataLoader workers (det3d/datasets/pipelines/pose.py:186-255, :385-452;
//       det3d/core/utils/center_utils.py:67-91);
//   N2  the optimizer step on the flat buffers: global L2-norm clip (hooks/optimizer.py:9-24, max_norm 35) + decoupled
//       weight decay (solver/fastai_optim.py:158-174, true_wd, bn_wd=True) + Adam (torch.optim.Adam semantics) in one
//       pass over (param, grad, m, v) — instead of ~155 x 4 tiny framework kernels.
// Both are HBM-bound streaming kernels.
#include "common.cuh"

namespace {

constexpr int kNormBlocks = 1024;

// ------------------------------------------------------------------------------------------------ N2
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float sh[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sh[w];
    partial[blockIdx.x] = s;
  }
}

// hyper (optional, device): {lr, beta1, beta2, eps, wd, bias1, bias2_sqrt, max_norm} overriding the by-value arguments,
// so a captured CUDA graph of the step can be replayed with a new schedule point without re-capturing
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float lr, float beta1, float beta2,
                                                        float eps, float wd, float bias1, float bias2_sqrt, float max_norm,
                                                        const float* __restrict__ hyper, const float* __restrict__ partial,
                                                        float* __restrict__ norm_out) {
  __shared__ float s_coef;
  __shared__ double s_tot[256];
  if (hyper) {
    lr = hyper[0]; beta1 = hyper[1]; beta2 = hyper[2]; eps = hyper[3]; wd = hyper[4]; bias1 = hyper[5]; bias2_sqrt = hyper[6];
    max_norm = hyper[7];
  }
  // every block reduces the kNormBlocks partial sums itself, in a fixed order (strided per thread, then a tree): deterministic,
  // identical in all blocks (one thread adding them one by one put ~10 us in front of every block's update loop)
  {
    double t = 0;
    for (int i = threadIdx.x; i < kNormBlocks; i += 256) t += partial[i];
    s_tot[threadIdx.x] = t;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) s_tot[threadIdx.x] += s_tot[threadIdx.x + s];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const double tot = s_tot[0];
    const float norm = (float)sqrt(tot);
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;  // torch.nn.utils.clip_grad_norm_
    s_coef = coef < 1.f ? coef : 1.f;
    if (blockIdx.x == 0 && norm_out) *norm_out = norm;
  }
  __syncthreads();
  const float coef = s_coef, decay = 1.f - wd * lr, step = lr / bias1;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float pi = p[i] * decay;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step * mi / (sqrtf(vi) / bias2_sqrt + eps);
  }
}

// ------------------------------------------------------------------------------------------------ N1
struct Tgt {
  const double* poses;  // [B][15][3] metres (x, y, z), float64 like the reference's python floats
  float* hm;           // [B][ncls][Z][Y][X], zero-filled by the caller kernel below
  int64_t* ind;
  uint8_t* mask;
  int64_t* cat;
  float* anno;         // [B][M][R]
  int B, Z, Y, X, one_hm, radius;
  double vx, vy, vz;   // voxel size
  float x0, y0, z0;    // range minima, float32 as in the reference (pose.py:190,389)
};

__global__ void zero_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// one block per (frame, heat-map entry): entry = pelvis only (one_hm) or each of the 15 joints
__global__ void __launch_bounds__(128) assign_targets_kernel(Tgt t) {
  const int b = blockIdx.x, k = blockIdx.y;
  const int M = t.one_hm ? 1 : 15, R = t.one_hm ? 45 : 3, ncls = t.one_hm ? 1 : 15;
  const double* pose = t.poses + (int64_t)b * 45;
  // voxel-unit coordinates: python float (double) arithmetic against float32 range minima, rounded to float32
  auto coord = [&](int joint, int axis) -> float {
    const double p = pose[joint * 3 + axis];
    const double lo = axis == 0 ? (double)t.x0 : (axis == 1 ? (double)t.y0 : (double)t.z0);
    const double vs = axis == 0 ? t.vx : (axis == 1 ? t.vy : t.vz);
    return (float)((p - lo) / vs);
  };
  const float cxf = coord(k, 0), cyf = coord(k, 1), czf = coord(k, 2);
  const int cx = (int)cxf, cy = (int)cyf, cz = (int)czf;  // astype(int32): truncation
  const bool inside = cx >= 0 && cx < t.X && cy >= 0 && cy < t.Y && cz >= 0 && cz < t.Z;
  const int slot = b * M + k;
  if (threadIdx.x == 0) {
    t.ind[slot] = inside ? (int64_t)cz * t.Y * t.X + (int64_t)cy * t.X + cx : 0;
    t.mask[slot] = inside ? 1 : 0;
    t.cat[slot] = (inside && !t.one_hm) ? k : 0;
  }
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float a = 0.f;
    if (inside) {
      if (t.one_hm) {
        const int joint = r / 3, axis = r % 3;
        a = coord(joint, axis) - (float)(axis == 0 ? cx : (axis == 1 ? cy : cz));
      } else {
        a = (r == 0 ? cxf : (r == 1 ? cyf : czf)) - (float)(r == 0 ? cx : (r == 1 ? cy : cz));
      }
    }
    t.anno[(int64_t)slot * R + r] = a;
  }
  if (!inside) return;
  // gaussian3D (center_utils.py:67-72): exp(-(r^2) / (2 sigma^2)^(3/2)), sigma = (2r+1)/6; max-splat, clipped at the border
  const int rad = t.radius, d = 2 * rad + 1;
  const double sigma = (double)d / 6.0;
  const double den = pow(2.0 * sigma * sigma, 1.5);
  float* hm = t.hm + ((int64_t)b * ncls + (t.one_hm ? 0 : k)) * t.Z * t.Y * t.X;
  for (int i = threadIdx.x; i < d * d * d; i += blockDim.x) {
    const int dx = i % d - rad, dy = (i / d) % d - rad, dz = i / (d * d) - rad;
    const int x = cx + dx, y = cy + dy, z = cz + dz;
    if (x < 0 || x >= t.X || y < 0 || y >= t.Y || z < 0 || z >= t.Z) continue;
    const float g = (float)exp(-(double)(dx * dx + dy * dy + dz * dz) / den);
    // values are >= 0, so the float ordering equals the ordering of their bit patterns
    atomicMax(reinterpret_cast<int*>(hm + ((int64_t)z * t.Y + y) * t.X + x), __float_as_int(g));
  }
}

}  // namespace

extern "C" int64_t rtp_adam_workspace_bytes(void) { return (int64_t)kNormBlocks * 4; }

extern "C" int rtp_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, float wd, int32_t step, float max_norm, float* workspace, float* grad_norm_out,
                             void* stream) {
  RTP_CHECK_ARG(param && grad && m && v && workspace && n > 0 && step >= 1, "rtp_adam_step: bad arguments");
  RTP_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps > 0.f, "rtp_adam_step: bad hyper-parameters");
  sumsq_partial_kernel<<<kNormBlocks, 256, 0, (cudaStream_t)stream>>>(grad, n, workspace);
  const float bias1 = 1.f - powf(beta1, (float)step);
  const float bias2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const int blocks = ceil_div(n, 1024) > 592 ? 592 : ceil_div(n, 1024);
  adam_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, lr, beta1, beta2, eps, wd, bias1, bias2_sqrt,
                                                            max_norm, nullptr, workspace, grad_norm_out);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, const float* hyper, float* workspace,
                                 float* grad_norm_out, void* stream) {
  RTP_CHECK_ARG(param && grad && m && v && hyper && workspace && n > 0, "rtp_adam_step_dev: bad arguments");
  sumsq_partial_kernel<<<kNormBlocks, 256, 0, (cudaStream_t)stream>>>(grad, n, workspace);
  const int blocks = ceil_div(n, 1024) > 592 ? 592 : ceil_div(n, 1024);
  adam_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, 0.f, 0.f, 0.f, 1.f, 0.f, 1.f, 1.f, 0.f, hyper, workspace,
                                                            grad_norm_out);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_assign_targets(const double* poses, int32_t B, int32_t Z, int32_t Y, int32_t X, int32_t one_hm, int32_t radius,
                                  const double* voxel_xyz, const float* range_xyz, float* hm, int64_t* ind, uint8_t* mask,
                                  int64_t* cat, float* anno, void* stream) {
  RTP_CHECK_ARG(poses && voxel_xyz && range_xyz && hm && ind && mask && cat && anno, "rtp_assign_targets: null pointer");
  RTP_CHECK_ARG(B > 0 && Z > 0 && Y > 0 && X > 0 && radius >= 0 && radius <= 4, "rtp_assign_targets: bad sizes");
  Tgt t{poses, hm, ind, mask, cat, anno, B, Z, Y, X, one_hm ? 1 : 0, radius, voxel_xyz[0], voxel_xyz[1], voxel_xyz[2],
        range_xyz[0], range_xyz[1], range_xyz[2]};
  const int64_t nhm = (int64_t)B * (one_hm ? 1 : 15) * Z * Y * X;
  zero_kernel<<<ceil_div(nhm, 1024) > 1184 ? 1184 : ceil_div(nhm, 1024), 256, 0, (cudaStream_t)stream>>>(hm, nhm);
  assign_targets_kernel<<<dim3(B, one_hm ? 1 : 15), 128, 0, (cudaStream_t)stream>>>(t);
  RTP_LAUNCH_CHECK();
}
