// head.cu — CenterHead loss (focal + masked L1, forward and gradient in one pass over the heatmap) and the
// arg-max keypoint decode.  HBM-bound: each reads the heatmap once.
//
// Algorithmic bytes per frame: loss = hm bf16 (2 B x 8-channel vector per voxel) + target fp32 (4 B per class
// voxel) read, d_hm written (+ a zero-filled K-padding chunk when the caller hands one over), d_reg zero-filled unless the
// caller consumes it at the target voxels only (RTP_LOSS_SPARSE_DREG); decode = hm read once + one regression row.
#include "common.cuh"

namespace {

constexpr float kPMin = 1e-4f, kPMax = 1.0f - 1e-4f;
constexpr int kRegStage = 3072;  // (target, regression channel) pairs of one round of samples in head_loss_final_kernel

struct LossK {
  P8 hm, reg, d_hm, d_reg;
  int ncls, R, M;
  const float* tgt;  // [N][ncls][Z][Y][X]
  const int64_t* ind;
  const uint8_t* mask;
  const int64_t* cat;
  const float* anno;
  float weight, grad_scale;
  const float* code_w;
  float* out;
  float* partial;  // [nblocks] neg-loss partial sums
  int has_grad;
  int sparse_dreg;  // RTP_LOSS_SPARSE_DREG: d_reg is written at the target voxels only (no dense zero-fill)
};

// number of set mask entries, counted by the whole block (every thread gets the total; exact in fp32 for < 2^24 targets)
__device__ __forceinline__ float count_mask(const uint8_t* __restrict__ mask, int n) {
  int total = 0;
  for (int base = 0; base < n; base += (int)blockDim.x) {
    const int i = base + (int)threadIdx.x;
    total += __syncthreads_count(i < n && mask[i] != 0);
  }
  return (float)total;
}

__device__ __forceinline__ float sigmoidf_(float h) { return 1.0f / (1.0f + expf(-h)); }

// One block per (32x32 (x,y) tile, z, n): target tiles are read x-fastest (coalesced in NCDHW) through smem,
// the P8 heatmap / gradients are accessed y-fastest.
__global__ void __launch_bounds__(256, 4) focal_neg_kernel(const __grid_constant__ LossK p) {
  __shared__ float tile[32][33];
  __shared__ float red[8];
  const P8& hm = p.hm;
  const int xt = blockIdx.x * 32, yt = blockIdx.y * 32;
  const int z = blockIdx.z % hm.Z, n = blockIdx.z / hm.Z;
  const int tid = threadIdx.x;
  const float numpos = count_mask(p.mask, hm.N * p.M);
  const float inv_np = numpos > 0.f ? 1.0f / numpos : 1.0f;
  const int64_t vol = (int64_t)hm.Z * hm.Y * hm.X;
  float neg = 0.f;
  const int nch = (p.ncls + 7) / 8;
  for (int ch = 0; ch < nch; ++ch) {
    // this thread's 4 voxels x 8 classes: logits and gradients stay packed (2 bf16 per register) — with fp32 arrays the
    // kernel needed 114 registers, i.e. 16 resident warps per SM for a pass that is pure memory latency
    uint4 hv[4], gv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = tid + k * 256;
      const int yy = i & 31, xx = i >> 5;
      const int x = xt + xx, y = yt + yy;
      hv[k] = (x < hm.X && y < hm.Y) ? ldg16(hm.ptr + n * hm.n_stride + ch * hm.c_stride + hm.voxel(z, x, y)) : make_uint4(0, 0, 0, 0);
      gv[k] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (ch * 8 + c < p.ncls) {  // block-uniform
        const float* tg = p.tgt + ((int64_t)n * p.ncls + ch * 8 + c) * vol + (int64_t)z * hm.Y * hm.X;
        __syncthreads();
        for (int i = tid; i < 1024; i += 256) {
          const int xx = i & 31, yy = i >> 5;
          const int x = xt + xx, y = yt + yy;
          tile[yy][xx] = (x < hm.X && y < hm.Y) ? tg[(int64_t)y * hm.X + x] : 1.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = tid + k * 256;
          const int yy = i & 31, xx = i >> 5;
          const int x = xt + xx, y = yt + yy;
          if (x < hm.X && y < hm.Y) {
            const uint32_t hw = c < 2 ? hv[k].x : c < 4 ? hv[k].y : c < 6 ? hv[k].z : hv[k].w;
            const float h = __uint_as_float((c & 1) ? (hw & 0xffff0000u) : (hw << 16));
            const float t = tile[yy][xx];
            const float s = sigmoidf_(h);
            const float pr = fminf(fmaxf(s, kPMin), kPMax);
            const float omt = 1.f - t;
            const float gt = omt * omt * omt * omt;
            const float l1p = logf(1.f - pr);
            neg += l1p * pr * pr * gt;
            // d/dh of -(neg)/num_pos ; clamp passes gradient only inside [kPMin, kPMax]
            const bool inside = s >= kPMin && s <= kPMax;
            const float dneg_dp = (2.f * pr * l1p - pr * pr / (1.f - pr)) * gt;
            const float g = inside ? -inv_np * dneg_dp * pr * (1.f - pr) * p.grad_scale : 0.f;
            const uint32_t gb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(g)) << ((c & 1) * 16);
            if (c < 2) gv[k].x |= gb; else if (c < 4) gv[k].y |= gb; else if (c < 6) gv[k].z |= gb; else gv[k].w |= gb;
          }
        }
      }
    }
    if (p.has_grad) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        const int yy = i & 31, xx = i >> 5;
        const int x = xt + xx, y = yt + yy;
        if (x < hm.X && y < hm.Y)
          stg16(p.d_hm.ptr + n * p.d_hm.n_stride + ch * p.d_hm.c_stride + p.d_hm.voxel(z, x, y), gv[k]);
      }
    }
  }
  if (p.has_grad) {
    // d_reg is sparse: zero-fill here (unless the caller consumes it at the target voxels only), the target voxels are
    // written by the final kernel.  Chunks of d_hm behind the class chunks (a spare K-padding chunk, Engine.loss) are
    // zero-filled too.
    const uint4 zero = make_uint4(0, 0, 0, 0);
    const int rch = p.sparse_dreg ? 0 : (p.R + 7) / 8;
    for (int k = 0; k < 4; ++k) {
      const int i = tid + k * 256;
      const int yy = i & 31, xx = i >> 5;
      const int x = xt + xx, y = yt + yy;
      if (x < hm.X && y < hm.Y) {
        for (int c = 0; c < rch; ++c)
          stg16(p.d_reg.ptr + n * p.d_reg.n_stride + c * p.d_reg.c_stride + p.d_reg.voxel(z, x, y), zero);
        for (int c = nch; c < p.d_hm.C8; ++c)
          stg16(p.d_hm.ptr + n * p.d_hm.n_stride + c * p.d_hm.c_stride + p.d_hm.voxel(z, x, y), zero);
      }
    }
  }
  neg = warp_sum(neg);
  if ((tid & 31) == 0) red[tid >> 5] = neg;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    p.partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
  }
}

// single block: reduce neg partials (fixed order), positive-location terms, regression L1, gradients at `ind`
__global__ void __launch_bounds__(256) head_loss_final_kernel(const __grid_constant__ LossK p, int npartial) {
  __shared__ double sred[256];
  __shared__ float s_reg[64];
  const int tid = threadIdx.x;
  const P8& hm = p.hm;
  double acc = 0;
  for (int i = tid; i < npartial; i += 256) acc += p.partial[i];
  sred[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) sred[tid] += sred[tid + s];
    __syncthreads();
  }
  // The per-target work is spread over the block (thread = target for the heat-map terms, thread = regression channel
  // for the L1 terms) but every sum keeps the serial order of the reference loop, so the result does not depend on
  // the thread count.  (One thread walking all N*M*R dependent global loads took 0.22 ms.)
  __shared__ float s_pos[256];
  const int NM = hm.N * p.M;
  const int YX = hm.Y * hm.X;
  const float num_pos = count_mask(p.mask, NM);
  const float inv_np = num_pos > 0.f ? 1.f / num_pos : 1.f;
  const float reg_den = num_pos + 1e-4f;
  // heat-map positive terms: target i -> s_pos (chunks of 256 targets, summed in order by thread 0)
  float pos = 0.f;
  for (int base = 0; base < NM; base += 256) {
    const int i = base + tid;
    float term = 0.f;
    if (i < NM) {
      const int n = i / p.M;
      const float m = p.mask[i] ? 1.f : 0.f;
      const int64_t id = p.ind[i];
      const int z = (int)(id / YX), y = (int)((id % YX) / hm.X), x = (int)(id % hm.X);
      const int c = (int)p.cat[i];
      const bf16* hp = hm.ptr + n * hm.n_stride + (c >> 3) * hm.c_stride + hm.voxel(z, x, y) + (c & 7);
      const float sg = sigmoidf_(__bfloat162float(*hp));
      const float pr = fminf(fmaxf(sg, kPMin), kPMax);
      const float lp = logf(pr);
      term = lp * (1.f - pr) * (1.f - pr) * m;
      if (p.has_grad && m > 0.f && num_pos > 0.f && sg >= kPMin && sg <= kPMax) {
        // two targets of one sample may share a voxel only with different classes (cat), i.e. different elements
        const float dpos_dp = (1.f - pr) * (1.f - pr) / pr - 2.f * (1.f - pr) * lp;
        bf16* gp = p.d_hm.ptr + n * p.d_hm.n_stride + (c >> 3) * p.d_hm.c_stride + p.d_hm.voxel(z, x, y) + (c & 7);
        const float g = __bfloat162float(*gp) - inv_np * dpos_dp * pr * (1.f - pr) * p.grad_scale;
        *gp = __float2bfloat16(g);
      }
    }
    s_pos[tid] = term;
    __syncthreads();
    if (tid == 0)
      for (int j = 0; j < 256 && base + j < NM; ++j) pos += s_pos[j];
    __syncthreads();
  }
  // regression: the |pred - target| terms and gradient contributions of all (target, channel) pairs are evaluated in
  // parallel, a block of samples at a time; thread r then adds the terms of channel r in target order (the reference's
  // order), and a target's gradient element is the bf16 accumulation chain over the sample's targets on that voxel, in
  // order (what a read-modify-write per target produces; one thread walking the targets paid three dependent loads each).
  __shared__ float s_term[kRegStage], s_g[kRegStage];
  __shared__ int s_id[kRegStage];
  const int MR = p.M * p.R;
  const int spc = max(1, kRegStage / MR);  // samples per round (the launcher checks M * R <= kRegStage)
  float sr = 0.f;
  for (int n0 = 0; n0 < hm.N; n0 += spc) {
    const int ns = min(spc, hm.N - n0), npair = ns * MR;
    for (int j = tid; j < ns * p.M; j += 256) s_id[j] = (int)p.ind[n0 * p.M + j];
    for (int e = tid; e < npair; e += 256) {
      const int li = e / p.R, r = e - li * p.R, i = n0 * p.M + li, n = i / p.M;
      const float m = p.mask[i] ? 1.f : 0.f;
      const int64_t id = p.ind[i];
      const int z = (int)(id / YX), y = (int)((id % YX) / hm.X), x = (int)(id % hm.X);
      const bf16* rp = p.reg.ptr + n * p.reg.n_stride + (r >> 3) * p.reg.c_stride + p.reg.voxel(z, x, y) + (r & 7);
      const float pred = __bfloat162float(*rp) * m;
      const float tg = p.anno[(int64_t)i * p.R + r] * m;
      const float diff = pred - tg;
      s_term[e] = fabsf(diff) / reg_den;
      const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      s_g[e] = m > 0.f ? p.weight * p.code_w[r] * sgn / reg_den * p.grad_scale : 0.f;
    }
    __syncthreads();
    if (p.has_grad)
      for (int e = tid; e < npair; e += 256) {
        const int li = e / p.R, r = e - li * p.R, n = n0 + li / p.M, l0 = (li / p.M) * p.M;
        const int id = s_id[li];
        float v = 0.f;  // d_reg is zero at the target voxels before this kernel (zero-filled) or not yet written (sparse mode)
        for (int jj = 0; jj < p.M; ++jj)
          if (s_id[l0 + jj] == id) v = __bfloat162float(__float2bfloat16(v + s_g[(l0 + jj) * p.R + r]));
        const int z = id / YX, y = (id % YX) / hm.X, x = id % hm.X;
        p.d_reg.ptr[n * p.d_reg.n_stride + (r >> 3) * p.d_reg.c_stride + p.d_reg.voxel(z, x, y) + (r & 7)] = __float2bfloat16(v);
      }
    if (tid < p.R)
      for (int li = 0; li < ns * p.M; ++li) sr += s_term[li * p.R + tid];
    __syncthreads();
  }
  if (tid < p.R) s_reg[tid] = sr;
  __syncthreads();
  if (tid == 0) {
    const float neg = (float)sred[0];
    const float hm_loss = num_pos > 0.f ? -(pos + neg) / num_pos : -neg;
    float loc = 0.f;
    for (int r = 0; r < p.R; ++r) {
      loc += s_reg[r] * p.code_w[r];
      p.out[4 + r] = s_reg[r];
    }
    p.out[0] = hm_loss + p.weight * loc;
    p.out[1] = hm_loss;
    p.out[2] = loc;
    p.out[3] = num_pos;
  }
}

// ---------------------------------------------------------------------------------------------- decode
struct DecodeK {
  P8 hm, reg;
  int ncls, R;
  float vx, vy, vz, x0, y0, z0;
  int* out_index;
  float* out_score;
  float* out_xyz;
};

// one block per (sample, 8-class chunk): running (max logit, lowest reference flat index) per class
__global__ void __launch_bounds__(512) decode_kernel(const __grid_constant__ DecodeK p) {
  __shared__ float s_val[16][8];
  __shared__ int s_idx[16][8];
  const P8& hm = p.hm;
  const int ch = blockIdx.x, n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t V = (int64_t)hm.Z * hm.X * hm.Y;
  float best[8];
  int bidx[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { best[c] = -INFINITY; bidx[c] = 0x7fffffff; }
  const bf16* base = hm.ptr + n * hm.n_stride + ch * hm.c_stride;
  for (int64_t v = tid; v < V; v += 512) {
    int64_t q = v;
    const int y = (int)(q % hm.Y); q /= hm.Y;
    const int x = (int)(q % hm.X);
    const int z = (int)(q / hm.X);
    const int ref = (z * hm.Y + y) * hm.X + x;  // the reference's flat index (z*Y*X + y*X + x)
    float f[8];
    unpack8(ldg16(base + hm.voxel(z, x, y)), f);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (f[c] > best[c] || (f[c] == best[c] && ref < bidx[c])) { best[c] = f[c]; bidx[c] = ref; }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best[c], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx[c], o);
      if (ov > best[c] || (ov == best[c] && oi < bidx[c])) { best[c] = ov; bidx[c] = oi; }
    }
    if (lane == 0) { s_val[warp][c] = best[c]; s_idx[warp][c] = bidx[c]; }
  }
  __syncthreads();
  if (tid < 8 && ch * 8 + tid < p.ncls) {
    const int c = tid, cls = ch * 8 + tid;
    float bv = s_val[0][c];
    int bi = s_idx[0][c];
    for (int w = 1; w < 16; ++w)
      if (s_val[w][c] > bv || (s_val[w][c] == bv && s_idx[w][c] < bi)) { bv = s_val[w][c]; bi = s_idx[w][c]; }
    p.out_index[n * p.ncls + cls] = bi;
    p.out_score[n * p.ncls + cls] = sigmoidf_(bv);
    const int YX = hm.Y * hm.X;
    const int z = bi / YX, y = (bi % YX) / hm.X, x = bi % hm.X;
    const bf16* rb = p.reg.ptr + n * p.reg.n_stride + p.reg.voxel(z, x, y);
    float* o = p.out_xyz + ((int64_t)n * p.ncls + cls) * p.R;
    for (int r = 0; r < p.R; r += 3) {
      const float rx = __bfloat162float(rb[(r >> 3) * p.reg.c_stride + (r & 7)]);
      const float ry = __bfloat162float(rb[((r + 1) >> 3) * p.reg.c_stride + ((r + 1) & 7)]);
      const float rz = __bfloat162float(rb[((r + 2) >> 3) * p.reg.c_stride + ((r + 2) & 7)]);
      // (idx + reg) * voxel + range, separately rounded mul and add as in the reference's tensor ops
      o[r] = __fadd_rn(__fmul_rn(__fadd_rn((float)x, rx), p.vx), p.x0);
      o[r + 1] = __fadd_rn(__fmul_rn(__fadd_rn((float)y, ry), p.vy), p.y0);
      o[r + 2] = __fadd_rn(__fmul_rn(__fadd_rn((float)z, rz), p.vz), p.z0);
    }
  }
}

}  // namespace

extern "C" int64_t rtp_head_loss_workspace_bytes(int32_t N, int32_t ncls, int32_t Z, int32_t Y, int32_t X) {
  (void)ncls;
  return (int64_t)N * Z * ceil_div(X, 32) * ceil_div(Y, 32) * 4;
}

extern "C" int rtp_head_loss_flags(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* tgt_hm, const int64_t* ind,
                                   const uint8_t* mask, const int64_t* cat, const float* anno, int32_t M, float weight,
                                   const float* code_weights, float grad_scale, float* out, rtp_p8 d_hm, rtp_p8 d_reg,
                                   int32_t flags, void* workspace, void* stream) {
  RTP_CHECK_ARG(hm.ptr && reg.ptr && tgt_hm && ind && mask && cat && anno && code_weights && out && workspace,
                "rtp_head_loss: null argument");
  RTP_CHECK_ARG(ncls >= 1 && ncls <= hm.C8 * 8 && R >= 1 && R <= 60 && R % 3 == 0 && R <= reg.C8 * 8 && M >= 1,
                "rtp_head_loss: bad ncls/R/M");
  RTP_CHECK_ARG((int64_t)M * R <= kRegStage && (int64_t)hm.Z * hm.Y * hm.X < (1ll << 31), "rtp_head_loss: M * R or the grid is too large");
  RTP_CHECK_ARG(hm.N == reg.N && hm.Z == reg.Z && hm.X == reg.X && hm.Y == reg.Y, "rtp_head_loss: hm/reg geometry mismatch");
  const bool has_grad = d_hm.ptr != nullptr;
  RTP_CHECK_ARG(has_grad == (d_reg.ptr != nullptr), "rtp_head_loss: d_hm and d_reg must both be given or both be NULL");
  LossK k;
  k.hm = P8(hm); k.reg = P8(reg); k.d_hm = P8(d_hm); k.d_reg = P8(d_reg);
  k.ncls = ncls; k.R = R; k.M = M; k.tgt = tgt_hm; k.ind = ind; k.mask = mask; k.cat = cat; k.anno = anno;
  k.weight = weight; k.grad_scale = grad_scale; k.code_w = code_weights; k.out = out; k.partial = (float*)workspace;
  k.has_grad = has_grad;
  k.sparse_dreg = (flags & RTP_LOSS_SPARSE_DREG) ? 1 : 0;
  if (has_grad) {
    RTP_CHECK_ARG(d_hm.C8 >= ceil_div(ncls, 8) && d_reg.C8 * 8 >= R, "rtp_head_loss: gradient tensors too narrow");
    RTP_CHECK_ARG(d_hm.N == hm.N && d_hm.Z == hm.Z && d_hm.X == hm.X && d_hm.Y == hm.Y && d_reg.N == hm.N && d_reg.Z == hm.Z &&
                      d_reg.X == hm.X && d_reg.Y == hm.Y,
                  "rtp_head_loss: gradient geometry mismatch");
  }
  dim3 grid(ceil_div(hm.X, 32), ceil_div(hm.Y, 32), hm.N * hm.Z);
  focal_neg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k);
  head_loss_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(k, (int)(grid.x * grid.y * grid.z));
  RTP_LAUNCH_CHECK();
}

extern "C" int rtp_head_loss(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* tgt_hm, const int64_t* ind,
                             const uint8_t* mask, const int64_t* cat, const float* anno, int32_t M, float weight,
                             const float* code_weights, float grad_scale, float* out, rtp_p8 d_hm, rtp_p8 d_reg,
                             void* workspace, void* stream) {
  if (d_hm.ptr) d_hm.C8 = (ncls + 7) / 8;  // this entry point leaves chunks behind the class chunks alone
  return rtp_head_loss_flags(hm, reg, ncls, R, tgt_hm, ind, mask, cat, anno, M, weight, code_weights, grad_scale, out, d_hm, d_reg,
                             0, workspace, stream);
}

extern "C" int rtp_decode(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* voxel_xyz, const float* range_xyz,
                          int32_t* out_index, float* out_score, float* out_xyz, void* stream) {
  RTP_CHECK_ARG(hm.ptr && reg.ptr && voxel_xyz && range_xyz && out_index && out_score && out_xyz, "rtp_decode: null argument");
  RTP_CHECK_ARG(ncls >= 1 && ncls <= hm.C8 * 8 && R >= 3 && R % 3 == 0 && R <= reg.C8 * 8, "rtp_decode: bad ncls/R");
  RTP_CHECK_ARG(hm.N == reg.N && hm.Z == reg.Z && hm.X == reg.X && hm.Y == reg.Y, "rtp_decode: hm/reg geometry mismatch");
  DecodeK k;
  k.hm = P8(hm); k.reg = P8(reg); k.ncls = ncls; k.R = R;
  // voxel_xyz / range_xyz are HOST pointers (6 floats of configuration)
  k.vx = voxel_xyz[0]; k.vy = voxel_xyz[1]; k.vz = voxel_xyz[2];
  k.x0 = range_xyz[0]; k.y0 = range_xyz[1]; k.z0 = range_xyz[2];
  k.out_index = out_index; k.out_score = out_score; k.out_xyz = out_xyz;
  decode_kernel<<<dim3(ceil_div(ncls, 8), hm.N), 512, 0, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}
