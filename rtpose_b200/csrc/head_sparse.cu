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
constexpr int kWgStage = 512;  // target voxels staged per round in reg_wgrad_sparse_kernel

// block (i, n): the 27 neighbours u of target voxel i of sample n, all Cin channels
__global__ void __launch_bounds__(256) reg_dgrad_sparse_kernel(P8 dy, P8 tin, const int64_t* __restrict__ ind, int M, const float* __restrict__ w,
                                                               int R, int Cin, P8 dt, uint8_t* __restrict__ uniq) {
  __shared__ int s_z[kMaxM], s_y[kMaxM], s_x[kMaxM];
  __shared__ float s_dy[kMaxM][kMaxR];
  __shared__ int s_first[kMaxM];
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
  // s_first[j]: target j is the first of the sample on its voxel (dy at a shared voxel already holds the summed gradient)
  for (int j = tid; j < M; j += 256) {
    bool f = true;
    for (int k = 0; k < j; ++k) f = f && !(s_z[k] == s_z[j] && s_y[k] == s_y[j] && s_x[k] == s_x[j]);
    s_first[j] = f ? 1 : 0;
  }
  __syncthreads();
  const int wz = s_z[i], wy = s_y[i], wx = s_x[i];
  for (int e = tid; e < 27 * Cin; e += 256) {
    // e = ci * 27 + c, and the loop runs over the TARGETS (uniform across the warp), each thread with its own tap
    // t = (u - v_j) + 1 per axis: for the block's own target t == c, so a warp reads consecutive weights W[co][ci][t] and
    // all its lanes are active (a loop over the taps left one or two lanes active per iteration: 0.22 ms for 16 targets).
    const int ci = e / 27, c = e - ci * 27;
    const int uz = wz + c / 9 - 1, uy = wy + (c / 3) % 3 - 1, ux = wx + c % 3 - 1;
    if (uz < 0 || uz >= dy.Z || uy < 0 || uy >= dy.Y || ux < 0 || ux >= dy.X) continue;
    float acc = 0.f;
    for (int j = 0; j < M; ++j) {  // y[v] = sum_t W[t] x[v + o_t]  =>  dx[u] = sum_{targets v} W[t : o_t = u - v] dy[v]
      if (!s_first[j]) continue;
      const int dz = uz - s_z[j], dyy = uy - s_y[j], dx = ux - s_x[j];
      if (dz < -1 || dz > 1 || dyy < -1 || dyy > 1 || dx < -1 || dx > 1) continue;
      const int t = (dz + 1) * 9 + (dyy + 1) * 3 + (dx + 1);
      const float* wp = w + (int64_t)ci * 27 + t;
      for (int co = 0; co < R; ++co) acc = fmaf(wp[(int64_t)co * Cin * 27], s_dy[j][co], acc);
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
  // the unique target voxels are staged in shared memory, a round of kWgStage targets at a time: the loop below then has no
  // dependent index loads in front of its two operand loads
  __shared__ int s_n[kWgStage], s_z[kWgStage], s_y[kWgStage], s_x[kWgStage];
  const int e = blockIdx.x * 256 + threadIdx.x;
  const bool live = e < R * Cin * 27;
  const int t = e % 27, r = e / 27;
  const int ci = live ? r % Cin : 0, co = live ? r / Cin : 0;
  const int oz = t / 9 - 1, oy = (t / 3) % 3 - 1, ox = t % 3 - 1;
  float acc = 0.f, bsum = 0.f;
  for (int k0 = 0; k0 < N * M; k0 += kWgStage) {
    const int nk = min(kWgStage, N * M - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < nk; k += 256) {
      s_n[k] = uniq[k0 + k] ? (k0 + k) / M : -1;
      coords(ind[k0 + k], dy.Y, dy.X, s_z[k], s_y[k], s_x[k]);
    }
    __syncthreads();
    if (!live) continue;
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
      const int n = s_n[k];
      if (n < 0) continue;
      const int z = s_z[k], y = s_y[k], x = s_x[k];
      const float g = __bfloat162float(dy.ptr[(int64_t)n * dy.n_stride + (int64_t)(co >> 3) * dy.c_stride + dy.voxel(z, x, y) + (co & 7)]);
      bsum += g;
      const int iz = z + oz, iy = y + oy, ix = x + ox;
      if (iz < 0 || iz >= dy.Z || iy < 0 || iy >= dy.Y || ix < 0 || ix >= dy.X) continue;
      const float v = __bfloat162float(tin.ptr[(int64_t)n * tin.n_stride + (int64_t)(ci >> 3) * tin.c_stride + tin.voxel(iz, ix, iy) + (ci & 7)]);
      acc = fmaf(g, v, acc);
    }
  }
  if (!live) return;
  dW[e] = acc_w ? dW[e] + acc : acc;
  if (ci == 0 && t == 0) db[co] = acc_b ? db[co] + bsum : bsum;
}

}  // namespace

extern "C" int64_t rtp_reg_head_bwd_sparse_workspace_bytes(int32_t N, int32_t M) { return (int64_t)N * M + 16; }

namespace {
void launch_zero_chunks(const P8& t, int C8, cudaStream_t stream) {
  // max-shared L1 split: the fill can then start on SMs whose split is pinned by resident tensor-core CTAs (it is issued beside
  // the head's convolutions, Engine.head) instead of waiting for them to leave
  static bool configured_dev[RTP_MAX_DEVICES];
  bool& configured = configured_dev[rtp_current_device()];
  if (!configured) {
    cudaFuncSetAttribute(zero_chunks_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  const int64_t nvec = (int64_t)t.Z * t.Xp * t.Yp;
  int zb = (int)((nvec + 2047) / 2048);
  if (zb > 64) zb = 64;
  zero_chunks_kernel<<<dim3(zb, C8, t.N), 256, 0, stream>>>(t, C8);
}

int reg_head_bwd_sparse_impl(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R, int32_t Cin, rtp_p8 dt,
                             float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b, void* workspace, bool prezeroed,
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
  if (!prezeroed) launch_zero_chunks(dtp, C8, (cudaStream_t)stream);
  reg_dgrad_sparse_kernel<<<dim3(M, dy.N), 256, 0, (cudaStream_t)stream>>>(dy, tin, ind, M, w, R, Cin, dtp, (uint8_t*)workspace);
  reg_wgrad_sparse_kernel<<<ceil_div((int64_t)R * Cin * 27, 256), 256, 0, (cudaStream_t)stream>>>(dy, tin, ind, (const uint8_t*)workspace, dy.N,
                                                                                                 M, R, Cin, dW, accumulate_w, db,
                                                                                                 accumulate_b);
  RTP_LAUNCH_CHECK();
}
}  // namespace

extern "C" int rtp_reg_head_bwd_sparse(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R, int32_t Cin,
                                       rtp_p8 dt, float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b, void* workspace,
                                       void* stream) {
  return reg_head_bwd_sparse_impl(d_reg, t_in, ind, M, w, R, Cin, dt, dW, accumulate_w, db, accumulate_b, workspace, false, stream);
}

extern "C" int rtp_reg_head_bwd_sparse_prezeroed(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R,
                                                 int32_t Cin, rtp_p8 dt, float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b,
                                                 void* workspace, void* stream) {
  return reg_head_bwd_sparse_impl(d_reg, t_in, ind, M, w, R, Cin, dt, dW, accumulate_w, db, accumulate_b, workspace, true, stream);
}

extern "C" int rtp_zero_chunks(rtp_p8 t, void* stream) {
  RTP_CHECK_ARG(t.ptr && t.N >= 1 && t.C8 >= 1, "rtp_zero_chunks: bad tensor");
  const P8 tp(t);
  RTP_CHECK_ARG(tp.c_stride >= (int64_t)tp.Z * tp.Xp * tp.Yp * 8, "rtp_zero_chunks: chunk volumes must be dense");
  launch_zero_chunks(tp, tp.C8, (cudaStream_t)stream);
  RTP_LAUNCH_CHECK();
}
