// eval_aux.cu — pose-error metrics on the device ("evaluation post-step"):
// PJPE / ABS_PJPE of eval_util.py:5-11 per frame and their per-sequence, per-joint means in millimetres as
// CRUW_POSE_Dataset.evaluation computes them (det3d/datasets/cruw_pose/cruw_pose.py:277-295).
// The reference works in float64 numpy on python floats taken from fp32 tensors; the kernels do the same arithmetic
// in fp64 with explicitly rounded operations (no FMA contraction) in numpy's order, so results agree to the last bit
// for the per-frame errors: norm = sqrt((d0*d0 + d1*d1) + d2*d2).
#include "common.cuh"

namespace {

__device__ __forceinline__ double norm3(double a, double b, double c) {
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
}

// one thread per (frame, joint)
__global__ void pjpe_kernel(const float* __restrict__ pred, const double* __restrict__ gt, int N, int J, double* __restrict__ rel,
                            double* __restrict__ ab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * J) return;
  const int n = i / J;
  const float* p = pred + (int64_t)i * 3;
  const double* g = gt + (int64_t)i * 3;
  const float* p0 = pred + (int64_t)n * J * 3;  // joint 0 = root (pred -= pred[:1], gt -= gt[:1]; eval_util.py:6-7)
  const double* g0 = gt + (int64_t)n * J * 3;
  double da[3], dr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    da[k] = __dsub_rn((double)p[k], g[k]);
    dr[k] = __dsub_rn(__dsub_rn((double)p[k], (double)p0[k]), __dsub_rn(g[k], g0[k]));
  }
  ab[i] = norm3(da[0], da[1], da[2]);
  rel[i] = norm3(dr[0], dr[1], dr[2]);
}

// one thread per (sequence, joint, metric): mean over that sequence's frames in frame order, x 1000 (cruw_pose.py:291-295)
__global__ void pjpe_seq_mean_kernel(const double* __restrict__ rel, const double* __restrict__ ab, const int32_t* __restrict__ seq,
                                     int N, int J, int S, double* __restrict__ rel_mm, double* __restrict__ ab_mm,
                                     int32_t* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * J * 2) return;
  const int which = i / (S * J), sj = i - which * S * J, s = sj / J, j = sj - s * J;
  const double* src = which ? ab : rel;
  double a = 0.0;
  int c = 0;
  for (int n = 0; n < N; ++n)
    if (__ldg(seq + n) == s) {
      a = __dadd_rn(a, src[(int64_t)n * J + j]);
      ++c;
    }
  const double m = c ? __dmul_rn(__ddiv_rn(a, (double)c), 1000.0) : 0.0;
  (which ? ab_mm : rel_mm)[sj] = m;
  if (which == 0 && j == 0) count[s] = c;
}

}  // namespace

extern "C" int rtp_pjpe(const float* pred_xyz, const double* gt_xyz, int32_t N, int32_t J, double* out_rel, double* out_abs,
                        void* stream) {
  RTP_CHECK_ARG(pred_xyz && gt_xyz && out_rel && out_abs, "rtp_pjpe: null pointer");
  RTP_CHECK_ARG(N > 0 && J > 0, "rtp_pjpe: bad sizes N=%d J=%d", N, J);
  pjpe_kernel<<<ceil_div((int64_t)N * J, 256), 256, 0, (cudaStream_t)stream>>>(pred_xyz, gt_xyz, N, J, out_rel, out_abs);
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_pjpe_seq_mean(const double* rel, const double* abs_, const int32_t* seq_index, int32_t N, int32_t J, int32_t S,
                                 double* out_rel_mm, double* out_abs_mm, int32_t* out_count, void* stream) {
  RTP_CHECK_ARG(rel && abs_ && seq_index && out_rel_mm && out_abs_mm && out_count, "rtp_pjpe_seq_mean: null pointer");
  RTP_CHECK_ARG(N > 0 && J > 0 && S > 0, "rtp_pjpe_seq_mean: bad sizes");
  pjpe_seq_mean_kernel<<<ceil_div((int64_t)S * J * 2, 128), 128, 0, (cudaStream_t)stream>>>(rel, abs_, seq_index, N, J, S, out_rel_mm,
                                                                                             out_abs_mm, out_count);
  RTP_LAUNCH_CHECK();
}
