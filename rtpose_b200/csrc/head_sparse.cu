// head_sparse.cu — backward of the regression branch's last conv (SepHead "reg": conv3x3x3 hidden -> R, center_head.py:66-109)
// from the SPARSE loss gradient.  CenterHead.loss gathers the regression output at the target voxels only
// (center_head.py:244-270, `ind` / `mask`), so dL/d(reg) is zero except at <= M voxels per sample.  The dense path ran a
// full-resolution plane-streaming dgrad (R -> hidden), a full weight gradient and a channel sum over that all-zero tensor:
// 0.26 + 0.19 + 0.06 ms per step, all of it at the serial start of the backward pass.  Here:
//   * dL/d(hidden)[u] = [hidden[u] > 0] * sum_t sum_co W[co][ci][t] * dy[u - o_t]  is evaluated only for the voxels u in the
//     3x3x3 neighbourhood of a target voxel (each such u gets its FULL sum over all target voxels next to it, so overlapping
//     neighbourhoods write identical values: no atomics, fixed order), after one zero-fill of the tensor;
//   * dW[co][ci][t] = sum over target voxels w of dy[w][co] * hidden[w + o_t][ci], db[co] = sum_w dy[w][co]: one thread per
//     weight element walks the (unique) target voxels in order.
// dy is read back from the dense gradient tensor the loss kernel wrote (bf16, duplicates of a voxel already summed), so the
// operands are bit-identical to the dense path's; only fp32 summation order differs.
#include "common.cuh"

namespace {

// dst[n][c8][:] = 0 for the first C8 chunks (whole chunk volumes: pads are zero anyway)
__global__ void __launch_bounds__(256) zero_chunks_kernel(P8 t, int C8) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  uint4* p = reinterpret_cast<uint4*>(t.ptr + (int64_t)n * t.n_stride + (int64_t)c8 * t.c_stride);
  const int64_t nvec = (int64_t)t.Z * t.Xp * t.Yp;  // 16-byte vectors per chunk volume
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * 256) p[i] = z;
}

__device__ __forceinline__ void coords(int64_t id, int Y, int X, int& z, int& y, int& x) {  // reference flat index z*Y*X + y*X + x
  const int YX = Y * X;
  z = (int)(id / YX);
  const int r = (int)(id - (int64_t)z * YX);
  y = r / X;
  x = r - y * X;
}

constexpr int kMaxM = 64;
constexpr int kMaxR = 64;

// block (i, n): the 27 neighbours u of target voxel i of sample n, all Cin channels
__global__ void __launch_bounds__(256) reg_dgrad_sparse_kernel(P8 dy, P8 tin, const int64_t* __restrict__ ind, int M, const float* __restrict__ w,
                                                               int R, int Cin, P8 dt, uint8_t* __restrict__ uniq) {
  __shared__ int s_z[kMaxM], s_y[kMaxM], s_x[kMaxM];
  __shared__ float s_dy[kMaxM][kMaxR];
  const int i = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  for (int j = tid; j < M; j += 256) coords(ind[(int64_t)n * M + j], dy.Y, dy.X, s_z[j], s_y[j], s_x[j]);
  __syncthreads();
  // first occurrence of this voxel among the sample's targets?  (later duplicates repeat the same work: skip them)
  bool first = true;
  for (int j = 0; j < i; ++j) first = first && !(s_z[j] == s_z[i] && s_y[j] == s_y[i] && s_x[j] == s_x[i]);
  if (tid == 0) uniq[(int64_t)n * M + i] = first ? 1 : 0;
  if (!first) return;
  for (int e = tid; e < M * R; e += 256) {
    const int j = e / R, co = e - j * R;
    s_dy[j][co] = __bfloat162float(dy.ptr[(int64_t)n * dy.n_stride + (int64_t)(co >> 3) * dy.c_stride + dy.voxel(s_z[j], s_x[j], s_y[j]) + (co & 7)]);
  }
  __syncthreads();
  const int wz = s_z[i], wy = s_y[i], wx = s_x[i];
  for (int e = tid; e < 27 * Cin; e += 256) {
    const int c = e / Cin, ci = e - c * Cin;
    const int uz = wz + c / 9 - 1, uy = wy + (c / 3) % 3 - 1, ux = wx + c % 3 - 1;
    if (uz < 0 || uz >= dy.Z || uy < 0 || uy >= dy.Y || ux < 0 || ux >= dy.X) continue;
    float acc = 0.f;
    for (int t = 0; t < 27; ++t) {  // y[v] = sum_t W[t] x[v + o_t]  =>  dx[u] = sum_t W[t] dy[u - o_t]
      const int vz = uz - (t / 9 - 1), vy = uy - ((t / 3) % 3 - 1), vx = ux - (t % 3 - 1);
      int hit = -1;
      for (int j = 0; j < M; ++j)
        if (hit < 0 && s_z[j] == vz && s_y[j] == vy && s_x[j] == vx) hit = j;
      if (hit < 0) continue;
      const float* wp = w + (int64_t)ci * 27 + t;
      for (int co = 0; co < R; ++co) acc = fmaf(wp[(int64_t)co * Cin * 27], s_dy[hit][co], acc);
    }
    const int64_t off = (int64_t)n * tin.n_stride + (int64_t)(ci >> 3) * tin.c_stride + tin.voxel(uz, ux, uy) + (ci & 7);
    const float gate = __bfloat162float(tin.ptr[off]) > 0.f ? 1.f : 0.f;
    dt.ptr[(int64_t)n * dt.n_stride + (int64_t)(ci >> 3) * dt.c_stride + dt.voxel(uz, ux, uy) + (ci & 7)] = __float2bfloat16(acc * gate);
  }
}

// thread = weight element (co, ci, t); also db[co] by the threads with ci == 0, t == 0
__global__ void __launch_bounds__(256) reg_wgrad_sparse_kernel(P8 dy, P8 tin, const int64_t* __restrict__ ind, const uint8_t* __restrict__ uniq,
                                                               int N, int M, int R, int Cin, float* __restrict__ dW, int acc_w,
                                                               float* __restrict__ db, int acc_b) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= R * Cin * 27) return;
  const int t = e % 27, r = e / 27;
  const int ci = r % Cin, co = r / Cin;
  const int oz = t / 9 - 1, oy = (t / 3) % 3 - 1, ox = t % 3 - 1;
  float acc = 0.f, bsum = 0.f;
  for (int n = 0; n < N; ++n)
    for (int j = 0; j < M; ++j) {
      if (!uniq[(int64_t)n * M + j]) continue;
      int z, y, x;
      coords(ind[(int64_t)n * M + j], dy.Y, dy.X, z, y, x);
      const float g = __bfloat162float(dy.ptr[(int64_t)n * dy.n_stride + (int64_t)(co >> 3) * dy.c_stride + dy.voxel(z, x, y) + (co & 7)]);
      bsum += g;
      const int iz = z + oz, iy = y + oy, ix = x + ox;
      if (iz < 0 || iz >= dy.Z || iy < 0 || iy >= dy.Y || ix < 0 || ix >= dy.X) continue;
      const float v = __bfloat162float(tin.ptr[(int64_t)n * tin.n_stride + (int64_t)(ci >> 3) * tin.c_stride + tin.voxel(iz, ix, iy) + (ci & 7)]);
      acc = fmaf(g, v, acc);
    }
  dW[e] = acc_w ? dW[e] + acc : acc;
  if (ci == 0 && t == 0) db[co] = acc_b ? db[co] + bsum : bsum;
}

}  // namespace

extern "C" int64_t rtp_reg_head_bwd_sparse_workspace_bytes(int32_t N, int32_t M) { return (int64_t)N * M + 16; }

extern "C" int rtp_reg_head_bwd_sparse(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R, int32_t Cin,
                                       rtp_p8 dt, float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b, void* workspace,
                                       void* stream) {
  RTP_CHECK_ARG(d_reg.ptr && t_in.ptr && ind && w && dt.ptr && dW && db && workspace, "rtp_reg_head_bwd_sparse: null argument");
  RTP_CHECK_ARG(M >= 1 && M <= kMaxM && R >= 1 && R <= kMaxR && R <= d_reg.C8 * 8 && Cin >= 1 && Cin <= t_in.C8 * 8 && Cin <= dt.C8 * 8,
                "rtp_reg_head_bwd_sparse: bad M / R / Cin");
  RTP_CHECK_ARG(d_reg.N == t_in.N && d_reg.N == dt.N && d_reg.Z == t_in.Z && d_reg.X == t_in.X && d_reg.Y == t_in.Y && dt.Z == t_in.Z &&
                    dt.X == t_in.X && dt.Y == t_in.Y,
                "rtp_reg_head_bwd_sparse: geometry mismatch");
  const P8 dy(d_reg), tin(t_in), dtp(dt);
  const int C8 = ceil_div(Cin, 8);
  const int64_t nvec = (int64_t)dtp.Z * dtp.Xp * dtp.Yp;
  RTP_CHECK_ARG(dtp.c_stride >= nvec * 8, "rtp_reg_head_bwd_sparse: dt chunk volumes must be dense");
  int zb = (int)((nvec + 2047) / 2048);
  if (zb > 64) zb = 64;
  zero_chunks_kernel<<<dim3(zb, C8, dtp.N), 256, 0, (cudaStream_t)stream>>>(dtp, C8);
  reg_dgrad_sparse_kernel<<<dim3(M, dy.N), 256, 0, (cudaStream_t)stream>>>(dy, tin, ind, M, w, R, Cin, dtp, (uint8_t*)workspace);
  reg_wgrad_sparse_kernel<<<ceil_div((int64_t)R * Cin * 27, 256), 256, 0, (cudaStream_t)stream>>>(dy, tin, ind, (const uint8_t*)workspace, dy.N,
                                                                                                 M, R, Cin, dW, accumulate_w, db,
                                                                                                 accumulate_b);
  RTP_LAUNCH_CHECK();
}
