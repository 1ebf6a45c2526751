/* rtpose_b200.h — C ABI of librtpose_b200.so: the B200 (sm_100a) kernels behind the RT-Pose HRRadarPose
 * forward/backward hot path.
 *
 * The reference (ipl-uw/RT-POSE) has no FFI of its own for this path: every op below is, in the reference, a
 * PyTorch/ATen call made from Python (nn.Conv3d, nn.GroupNorm, F.interpolate, ...), plus one pybind11 module
 * (det3d/ops/dcn).  Each entry point therefore cites the reference *call site* it replaces (paths relative to
 * the reference root).  The host side that binds these symbols is rtpose_b200/lib.py (ctypes); the
 * reference-side stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - no allocation inside: callers pass outputs and workspaces (`*_workspace_bytes` queries);
 *   - return 0 on success, a negative value for an argument error, a positive cudaError_t otherwise;
 *     rtp_last_error() returns a human-readable message for the last failure on the calling thread;
 *   - re-entrant per stream; nothing here synchronises the device.
 *
 * Activation layout "P8" (rtp_p8): bf16, 8-channel blocked, in-plane zero-padded, y fastest:
 *       T[n][c/8][z][x+1][y+1][c%8],  extents [N][C8][Z][X+2][Y+2][8]
 *   The pad ring (x=-1, x=X, y=-1, y=Y) is zero and is never written by any kernel, so a 3x3x3 tap is a pure
 *   linear shift inside a z-plane and zero padding comes for free; `ptr` addresses element (0,0,0,-1,-1,0).
 *   Buffers must carry >= RTP_GUARD_BYTES of readable memory before and after (tiles read a small halo).
 */
#ifndef RTPOSE_B200_H_
#define RTPOSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTP_GUARD_BYTES 8192
#define RTP_MAX_TAPS 27

typedef struct {
  void* ptr;        /* bf16 element (n=0, chunk=0, z=0, xp=0, yp=0) */
  int64_t n_stride; /* elements between samples */
  int64_t c_stride; /* elements between 8-channel chunks (normally Z*(X+2)*(Y+2)*8) */
  int32_t N, C8, Z, X, Y;
} rtp_p8;

const char* rtp_last_error(void);
int rtp_version(void);
/* 1 if a device with compute capability 10.x is current, else 0 (kernels are sm_100a-only). */
int rtp_device_ok(void);
/* Device-wide L1 / shared-memory preference for the current device (cudaDeviceSetCacheConfig): 1 = prefer shared memory, so
 * that the streaming kernels can become resident beside the large-shared-memory persistent kernels (an SM's split only changes
 * when it is idle); 0 = no preference.  Call once per device before launching (lib.setup_device). */
int rtp_set_shared_carveout(int32_t prefer_shared);

/* ---- layout conversion at the det3d API boundary -------------------------------------------------------
 * replaces: the NCDHW fp32 tensors that flow between reference modules (radar_pose_net.py:26-46). */
int rtp_pack_ncdhw(const float* src, rtp_p8 dst, int32_t C, void* stream);                 /* fp32 [N,C,Z,Y,X] -> P8 */
int rtp_unpack_ncdhw(rtp_p8 src, float* dst, int32_t C, int32_t accumulate, void* stream); /* P8 -> fp32 [N,C,Z,Y,X] */

/* ---- radar-cube ingest -----------------------------------------------------------------------------------
 * replaces: CRUW_POSE_Dataset.get_cube / get_cube_phase (det3d/datasets/cruw_pose/cruw_pose.py:167-194), the
 * channel packing in AssignLabelPose(2).__call__ (det3d/datasets/pipelines/pose.py:163-172) and
 * RadarFeatureNet.forward (det3d/models/readers/radar_encoder.py:15-17).
 * raw: fp16 [N][D][RZ][RY][RX] (D Doppler bins or 2*D re/im planes, become channels); crops the ROI
 * z[z0,z0+Z) y[y0,y0+Y) x[x0,x0+X), applies (v - norm_start) / norm_scale, clamps < 0 to 0 (skipped when
 * normalize == 0), writes P8 bf16.  If dst_f32 != NULL also writes the fp32 [N,D,Z,Y,X] tensor the reference
 * would hand to the model. */
int rtp_ingest_pack(const void* raw_f16, int32_t N, int32_t D, int32_t RZ, int32_t RY, int32_t RX, int32_t z0,
                    int32_t y0, int32_t x0, float norm_start, float norm_scale, int32_t normalize, rtp_p8 dst,
                    float* dst_f32, void* stream);

/* ---- on-disk cubes: ROI-only reads (host code, no device work) ---------------------------------------------
 * replaces: `np.load(<seq>/DZYX_npy_f16/<frame>.npy).astype(np.float32)` followed by the z/y crop in
 * CRUW_POSE_Dataset.get_cube (det3d/datasets/cruw_pose/cruw_pose.py:170, :176-181) and get_cube_phase (:189-192);
 * the DataLoader workers / pickle / blocking H2D around them (det3d/datasets/loader/build_loader.py:46-57).
 * A cube file is a numpy .npy (v1-v3) holding little-endian float16 in C order with shape [..., RZ, RY, RX]; all
 * leading dimensions (Doppler bins, re/im planes) are flattened to `lead` planes.  rtp_npy_read_roi_slab copies rows
 * z[z0,z0+Z) y[y0,y0+Y) with their FULL x extent into dst as fp16 [lead][Z][Y][RX] using `threads` reader threads
 * (pread; dst is typically pinned memory that is then copied to the device and handed to rtp_ingest_pack with
 * RZ=Z, RY=Y, z0=y0=0, which crops x, normalises, clamps and packs).  Errors (return < 0, message in
 * rtp_last_error): missing/short file, bad magic or version, dtype other than float16, Fortran order, fewer than 3
 * dimensions, ROI outside the cube, dst too small. */
typedef struct {
  int32_t ndim;
  int32_t elem_bytes;
  int32_t fortran_order;
  int32_t reserved_;
  int64_t shape[8];
  int64_t data_offset; /* byte offset of element 0 in the file */
  int64_t file_bytes;
  char descr[16];      /* numpy dtype string, e.g. "<f2" */
} rtp_npy_info;
int rtp_npy_probe(const char* path, rtp_npy_info* info);
int64_t rtp_npy_roi_slab_bytes(const rtp_npy_info* info, int32_t Z, int32_t Y); /* < 0 on bad arguments */
int rtp_npy_read_roi_slab(const char* path, int32_t z0, int32_t Z, int32_t y0, int32_t Y, void* dst, int64_t dst_bytes,
                          int32_t threads);

/* ---- weights ---------------------------------------------------------------------------------------------
 * replaces: nothing in the reference (cuDNN consumes [Cout,Cin,kz,ky,kx] fp32 directly); repacks an fp32
 * nn.Conv3d weight into the bf16 UMMA B-operand tiles.
 *   mode 0 (forward): dst[tap][KP/8][NP][8],  K = Cin (padded to KP), N = Cout (padded to NP)
 *   mode 1 (dgrad)  : dst[tap][KP/8][NP][8],  K = Cout (padded to KP), N = Cin (padded to NP)
 * ci0/ci_n select an input-channel slice of w (used to split the 192->128 final conv per branch). */
int rtp_weight_pack(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin, int32_t ntaps, int32_t ci0,
                    int32_t ci_n, int32_t KP, int32_t NP, int32_t mode, void* stream);

/* ---- implicit-GEMM conv3d on tcgen05 tensor cores ---------------------------------------------------------
 * replaces: nn.Conv3d forward and the cuDNN dgrad of it — call sites hr_util/common.py:40,114;
 * hr_util/hr3d.py:84,148,169,185,298,324; hrnet3d.py:20; pose_heads/center_head.py:86,91,205 — with the
 * bias / ReLU / residual-add that follow them (common.py:146-147, center_head.py:88) fused in the epilogue.
 *
 * Rows (GEMM M) enumerate a grid (n, rz, rx, ry); for tap t the A row is the input vector at
 * (rz*IS+tz[t], rx*IS+tx[t], ry*IS+ty[t]) (zero outside), the B tile is packed-weight tap wt[t]; the result
 * goes to output voxel (rz*OS+oz0, rx*OS+ox0, ry*OS+oy0).  Forward stride-s conv: IS=s, OS=1; dgrad of a
 * stride-2 conv: IS=1, OS=2, one launch per output parity class. */
typedef struct {
  rtp_p8 in, out;
  rtp_p8 res;        /* optional residual, geometry of `out` (ptr NULL = none) */
  rtp_p8 mask;       /* optional ReLU mask tensor, geometry of `out`: result *= (mask > 0) */
  const void* w;     /* packed weights (rtp_weight_pack) */
  const float* bias; /* [NP] fp32 or NULL */
  int32_t Cin;       /* K per tap, multiple of 16 (chunks beyond in.C8 read as zero) */
  int32_t NP;        /* GEMM N, multiple of 16, <= 256 */
  int32_t out_c8;    /* 8-channel chunks of the result to store (<= NP/8) */
  int32_t ntaps;
  int8_t tz[RTP_MAX_TAPS], tx[RTP_MAX_TAPS], ty[RTP_MAX_TAPS], wt[RTP_MAX_TAPS];
  int32_t RZ, RX, RY;
  int32_t IS, OS, oz0, ox0, oy0;
  int32_t relu;       /* apply ReLU last */
  int32_t accumulate; /* out += result (read-modify-write) */
} rtp_conv_desc;
int rtp_conv(const rtp_conv_desc* d, void* stream);
/* Several descriptors in ONE launch: descs[0..n) (n <= 8) must agree in everything but the tap list and the row grid /
 * offsets.  Used for the 8 output-parity classes of a stride-2 dgrad, which at low resolution are each a handful of
 * latency-bound CTAs. */
int rtp_conv_multi(const rtp_conv_desc* descs, int32_t n, void* stream);

/* Plane-streaming 3x3x3 stride-1 conv (the dominant shape): the z-taps are stacked into GEMM N
 * (N = 3*NPo), accumulators for all output planes stay resident in TMEM, input planes are streamed once
 * through shared memory by 1-D bulk async copies; same epilogue options as rtp_conv.
 * w: packed by rtp_weight_pack_k3s1.  Cin % 16 == 0 (Cin <= 32 or a multiple of 32), NPo in {16..80}; outputs are
 * processed in z-chunks of 512/NPo planes.  gn_sums is reserved (must be NULL in this version). */
int rtp_weight_pack_k3s1(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin, int32_t KP, int32_t NPo,
                         int32_t transpose_flip, void* stream);
/* The same pack for the window [ci0, ci0 + ci_n) of the weight's input channels (w is [Cout][Cin_total][3][3][3]): the
 * per-group packs of the space-to-depth dgrad, without materialising the slice. */
int rtp_weight_pack_k3s1_window(const float* w, void* dst_bf16, int32_t Cout, int32_t Cin_total, int32_t ci0, int32_t ci_n,
                                int32_t KP, int32_t NPo, int32_t transpose_flip, void* stream);
/* All weight packs of a training step in one launch (the optimizer rewrites every weight, so every pack is rebuilt each
 * step: ops.PackedWeights.refresh_async).  jobs_dev: device array; kind 0 = rtp_weight_pack (flag = mode, ntaps), kind 1 =
 * rtp_weight_pack_k3s1[_window] (NP = NPo, flag = transpose_flip); block0 / nblocks: the job's range of 256-thread blocks,
 * ascending and contiguous from 0; total_blocks = their sum. */
typedef struct {
  const void* w;
  void* dst;
  int32_t kind, Cout, Cin_total, ntaps, ci0, ci_n, KP, NP, flag, block0, nblocks, reserved;
} rtp_pack_job;
int rtp_weight_pack_batch(const rtp_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);
typedef struct {
  rtp_p8 in, out, res, mask;
  const void* w;
  const float* bias;
  int32_t Cin, NPo, out_c8;
  int32_t relu, accumulate;
  /* fused per-(sample, channel) statistics of the result as stored (bf16-rounded), out_c8 <= 4; 0 = off.
   *   1: sum v, sum v*v          -> mean / rstd of the GroupNorm that consumes `out` (common.py:92-96 'g' layers)
   *   2: sum v, sum v*stat_aux   -> the two reductions of GroupNorm backward when `out` is dL/d(GN output) and
   *                                 stat_aux the GN input
   * Partial sums go to stat_ws (>= rtp_conv_k3s1_stat_ws_bytes(N)); rtp_conv_k3s1_stat_finalize turns them into
   * `stats` / `red` in a fixed summation order.  Saves one full read of the tensor per GroupNorm (forward and backward). */
  int32_t stat_mode;
  rtp_p8 stat_aux;
  float* stat_ws;
  /* structurally sparse weights (space-to-depth stride-2 convs): K is split into tap_mask_groups equal channel groups and
   * bit t9 = kx*3+ky of tap_mask[g] says whether in-plane tap t9 of group g has any non-zero weight; all-zero taps are
   * skipped by the MMA issuer.  use_tap_mask = 0: dense. */
  int32_t use_tap_mask, tap_mask_groups;
  uint16_t tap_mask[8];
  void* debug; /* tools only (tools/dbg_k3s1.py): [grid][8] int64 cycle counters of the MMA warp; NULL otherwise */
  /* optional list of the (sample, 128-position tile) units to process (rtp_active_units): device int32 unit ids u = n * ntile +
   * tile and their count; units not listed are skipped — with accumulate = 1 that leaves the existing output untouched, which
   * is exact when the input is known to be zero around them (the regression half of the head gradient).  NULL = all units. */
  const int32_t* unit_list;
  const int32_t* unit_count;
} rtp_conv_k3s1_desc;
int rtp_conv_k3s1(const rtp_conv_k3s1_desc* d, void* stream);
/* Units (n, tile) — tile = 128 consecutive in-plane positions of the padded plane, the unit of rtp_conv_k3s1 / rtp_wgrad_k3s1 —
 * that contain a voxel within `radius` (in x and y; every z belongs to a unit) of one of the voxels ind[n][0..M) (reference flat
 * index z*Y*X + y*X + x).  unit_list: int32 [N * ntile] (ascending ids, first *unit_count valid), ntile = ceil(X*(Y+2)/128).
 * One small launch; deterministic order. */
int rtp_active_units(const int64_t* ind, int32_t N, int32_t M, int32_t Z, int32_t X, int32_t Y, int32_t radius, int32_t* unit_list,
                     int32_t* unit_count, void* stream);
/* dynamic shared memory the kernel would use for this shape, or -1 when the shape is not supported (callers
 * then use rtp_conv) */
int64_t rtp_conv_k3s1_smem_bytes(int32_t Cin, int32_t NPo, int32_t Z, int32_t X, int32_t Y);
/* fused statistics (rtp_conv_k3s1_desc.stat_mode): workspace size, number of CTAs the launch for this shape uses, and
 * the finalisation: mode 1 -> stats[N][G][2] = (mean, rstd) as rtp_gn_finalize; mode 2 -> red[N][C][2] as
 * rtp_gn_bwd_reduce (stats_in = the forward statistics of the GroupNorm). */
int64_t rtp_conv_k3s1_stat_ws_bytes(int32_t N);
int32_t rtp_conv_k3s1_num_ctas(int32_t Cin, int32_t NPo, int32_t N, int32_t Z, int32_t X, int32_t Y);
int rtp_conv_k3s1_stat_finalize(const float* stat_ws, int32_t nctas, int32_t mode, int32_t N, int32_t C, int32_t G,
                                int64_t voxels, float eps, const float* stats_in, float* out, void* stream);

/* Pointwise (1x1x1) conv, forward or dgrad (w packed by rtp_weight_pack mode 0 / 1 with ntaps = 1): a streaming GEMM
 * over the linear positions of the P8 tensor (bulk-copy staging, resident weights, ping-pong TMEM accumulators).
 * replaces: the fuse-layer 1x1 convs (hr_util/hr3d.py:147-157), final_conv (backbones/hrnet3d.py:20,41) and their
 * dgrads.  K % 16 == 0 (<= 256), NP % 16 == 0 (<= 256); mask/bias/relu/accumulate as in rtp_conv. */
int rtp_conv_pw_supported(int32_t K, int32_t NP);
int rtp_conv_pw(rtp_p8 in, rtp_p8 out, rtp_p8 mask, const void* w, const float* bias, int32_t K, int32_t NP,
                int32_t out_c8, int32_t relu, int32_t accumulate, void* stream);

/* ---- weight gradient ---------------------------------------------------------------------------------------
 * replaces: cuDNN wgrad inside autograd's convolution_backward for the same call sites.
 * dW[tap][ci][co] = sum_rows X[row+tap][ci] * dY[row][co]; rows and taps as in rtp_conv_desc (X plays `in`,
 * dY is indexed by the row grid directly with OS/oz0.. = 1/0).  Split-K over `nsplit` CTAs into a fp32
 * workspace, then a deterministic reduction writes dW in the reference's [Cout][Cin][taps] fp32 layout. */
typedef struct {
  rtp_p8 x, dy;
  int32_t Cin;   /* multiple of 8 */
  int32_t NP;    /* dY channels padded to a multiple of 16 (chunks beyond dy.C8 read as zero) */
  int32_t ntaps;
  int8_t tz[RTP_MAX_TAPS], tx[RTP_MAX_TAPS], ty[RTP_MAX_TAPS];
  int32_t RZ, RX, RY, IS;
  int32_t nsplit;
  int16_t tc[RTP_MAX_TAPS]; /* first 8-channel chunk of x the tap reads (0 = plain tensor; parity group * Cin/8 when x is
                               a space-to-depth view and the taps are its 27 (parity, offset) pairs) */
  float* workspace;
} rtp_wgrad_desc;
int64_t rtp_wgrad_workspace_bytes(int32_t Cin, int32_t NP, int32_t ntaps, int32_t nsplit);
int rtp_wgrad(const rtp_wgrad_desc* d, void* stream);
/* dW[co][ci0+ci][tap] (fp32, [co_n][Cin_total][ntaps]) = or += sum over splits of the partial for dY channel
 * n0+co, input channel ci (co < co_n, ci < ci_n).  Cin8/NP/ntaps/nsplit as given to rtp_wgrad. */
int rtp_wgrad_reduce(const float* workspace, int32_t nsplit, int32_t Cin8, int32_t NP, int32_t ntaps, float* dW,
                     int32_t Cin_total, int32_t co_n, int32_t n0, int32_t ci0, int32_t ci_n, int32_t accumulate,
                     void* stream);

/* Plane-streaming weight gradient for 3x3x3 stride-1 convs with exactly 32 input channels (x: 4 chunks).  The three
 * z-taps are stacked along GEMM M, the 9 in-plane taps are 9 TMEM-resident accumulators; one persistent CTA per SM
 * writes an fp32 partial [nsplit][9][128][NP] and rtp_wgrad_k3s1_reduce sums them in a fixed order into
 * dW[co][ci0+ci][kz][ky][kx] (co < co_n reads dY channel n0+co).  zero_page: >= rtp_wgrad_k3s1_zero_bytes(Y) bytes of
 * device zeros; workspace: >= rtp_wgrad_k3s1_workspace_bytes(NP, #SMs).  *nsplit_out receives the partial count. */
int rtp_wgrad_k3s1_supported(int32_t Cin, int32_t NP, int32_t Z, int32_t X, int32_t Y);
int64_t rtp_wgrad_k3s1_workspace_bytes(int32_t NP, int32_t nsm);
int64_t rtp_wgrad_k3s1_zero_bytes(int32_t Y);
int rtp_wgrad_k3s1(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                   void* stream);
/* The same over the listed units only (rtp_active_units; exact when dy is zero in every other unit). */
int rtp_wgrad_k3s1_units(rtp_p8 x, rtp_p8 dy, int32_t NP, const void* zero_page, float* workspace, int32_t* nsplit_out,
                         const int32_t* unit_list, const int32_t* unit_count, void* stream);
int rtp_wgrad_k3s1_reduce(const float* workspace, int32_t nsplit, int32_t NP, float* dW, int32_t Cin_total,
                          int32_t co_n, int32_t n0, int32_t ci0, int32_t accumulate, void* stream);

/* ---- GroupNorm ---------------------------------------------------------------------------------------------
 * replaces: nn.GroupNorm(8, C) forward/backward — hr_util/common.py:57; hr_util/hr3d.py:83,147,168,184,297,323;
 * center_head.py:85,204.  Statistics are fp32 per (sample, group); eps = 1e-5, biased variance.
 *   sums : [N][C][2] fp32 per-channel (sum, sum of squares) over real voxels;
 *   stats: [N][G][2] fp32 (mean, rstd). */
/* reductions are two-level with a fixed slab count (deterministic); workspace >= rtp_gn_workspace_bytes(N, C8) */
int64_t rtp_gn_workspace_bytes(int32_t N, int32_t C8);
int rtp_gn_sums(rtp_p8 x, int32_t C, float* sums, float* workspace, void* stream);
int rtp_gn_finalize(const float* sums, int32_t N, int32_t C, int32_t G, int64_t voxels, float eps, float* stats,
                    void* stream);
/* rtp_gn_sums + rtp_gn_finalize in two launches instead of three (same arithmetic, C <= 256). */
int rtp_gn_stats(rtp_p8 x, int32_t C, int32_t G, float eps, float* stats, float* workspace, void* stream);
/* y = bf16((x - mean) * rstd * gamma + beta), pads stay zero */
int rtp_gn_apply(rtp_p8 x, int32_t C, int32_t G, const float* stats, const float* gamma, const float* beta, rtp_p8 y,
                 void* stream);
/* backward, two steps.  red[N][C][2] = per-channel (sum dy, sum dy*xhat) */
int rtp_gn_bwd_reduce(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, float* red, float* workspace,
                      void* stream);
/* dgamma/dbeta (+)= sum_n red; dx (=|+=) [x > 0]? (rstd*(gamma*dy - (s1 + xhat*s2)/m) + add): the (x > 0) factor when x
 * is itself a ReLU output (the gradient of every tensor is kept w.r.t. its pre-ReLU value); `add` (ptr NULL = none, x's
 * geometry) is a second gradient into the same tensor — the residual / fuse-sum pass-through of autograd's
 * AddBackward (hr_util/common.py:146, hr3d.py:213-227) — folded into this pass. */
int rtp_gn_bwd_apply(rtp_p8 x, rtp_p8 dy, int32_t C, int32_t G, const float* stats, const float* red,
                     const float* gamma, float* dgamma, float* dbeta, int32_t accumulate_params, rtp_p8 dx,
                     int32_t accumulate_dx, int32_t relu_mask, rtp_p8 add, void* stream);
/* Space-to-depth ("s2d") variants for the stride-2 exchange convs (fuse_layers / transition, hr_util/hr3d.py:159-203,
 * :262-292).  The s2d view of a tensor with grid (Z, X, Y) (all even) and C8 chunks is a P8 tensor with grid
 * (Z/2, X/2, Y/2) and 8*C8 chunks: voxel (z, x, y), chunk c lives at voxel (z/2, x/2, y/2), chunk
 * ((z&1)*4 + (x&1)*2 + (y&1))*C8 + c.  A stride-2 3x3x3 pad-1 conv over the tensor equals a stride-1 one over the view
 * with the weights of rtp_weight_s2d_expand, so it runs on rtp_conv_k3s1 / rtp_wgrad_k3s1 instead of the gather kernels.
 * rtp_gn_apply_s2d writes the normalised tensor directly as that view; the *_bwd_*_s2d calls read dL/d(view). */
int rtp_gn_apply_s2d(rtp_p8 x, int32_t C, int32_t G, const float* stats, const float* gamma, const float* beta,
                     rtp_p8 y_s2d, void* stream);
int rtp_gn_bwd_reduce_s2d(rtp_p8 x, rtp_p8 dy_s2d, int32_t C, int32_t G, const float* stats, float* red, float* workspace,
                          void* stream);
int rtp_gn_bwd_apply_s2d(rtp_p8 x, rtp_p8 dy_s2d, int32_t C, int32_t G, const float* stats, const float* red,
                         const float* gamma, float* dgamma, float* dbeta, int32_t accumulate_params, rtp_p8 dx,
                         int32_t accumulate_dx, int32_t relu_mask, rtp_p8 add, void* stream);
/* w [Cout][Cin][3][3][3] fp32 -> w_s2d [Cout][8*Cin][3][3][3] fp32 (zero except the 27 matching (parity, offset) pairs),
 * and the transpose for the weight gradient: dw (=|+=) fold(dw_s2d). */
/* Weight gradient of a 1x1x1 conv as a streaming GEMM over P8 positions (csrc/wgrad_pw.cu): dW[co][ci0+ci] (= or +=)
 * sum dY[co] * X[ci], dW fp32 [co_n][Cin_total], x: >= Cin channels, dy: <= 128 channels, same grid, dense planes.
 * replaces: autograd's weight gradient of the 1x1 fuse convs and of final_conv (hr_util/hr3d.py:144-158, hrnet3d.py:20). */
int rtp_wgrad_pw_supported(int32_t Cin, int32_t Cout, int32_t Z, int32_t X, int32_t Y);
int64_t rtp_wgrad_pw_workspace_bytes(int32_t Cin, int32_t nsm);
int rtp_wgrad_pw(rtp_p8 x, rtp_p8 dy, int32_t Cin, const void* zero_page, float* workspace, int32_t* nsplit_out, void* stream);
int rtp_wgrad_pw_reduce(const float* workspace, int32_t nsplit, int32_t Cin, float* dW, int32_t Cin_total, int32_t co_n,
                        int32_t ci0, int32_t accumulate, void* stream);
/* The same with the conv's bias gradient for free: one more GEMM N chunk is fed from ones_page (device, 128 positions x 8
 * bf16 channels, channel 0 = 1.0), whose product is sum over positions of dY; rtp_wgrad_pw_bias_reduce writes dW as above and
 * dbias[co] (= or +=).  replaces: autograd's bias gradient of final_conv (hrnet3d.py:20) — a separate pass over dY before. */
int64_t rtp_wgrad_pw_bias_workspace_bytes(int32_t Cin, int32_t nsm);
int rtp_wgrad_pw_bias(rtp_p8 x, rtp_p8 dy, int32_t Cin, const void* zero_page, const void* ones_page, float* workspace,
                      int32_t* nsplit_out, void* stream);
int rtp_wgrad_pw_bias_reduce(const float* workspace, int32_t nsplit, int32_t Cin, float* dW, int32_t Cin_total, int32_t co_n,
                             int32_t ci0, int32_t accumulate, float* dbias, int32_t accumulate_bias, void* stream);
/* Weight gradient of the stride-2 conv straight from the view, plane-streaming (csrc/wgrad_s2d.cu): xs = s2d view of the
 * normalised input (8*Cin/8 chunks), dy = gradient of the conv output (same grid as the view), Cin = 32.  One persistent
 * CTA per SM writes an fp32 partial [nsplit][6][128][2*NP]; rtp_wgrad_s2d_reduce sums them in a fixed order into
 * dW[co][ci][kz][ky][kx] (= or +=).  zero_page: >= 2048 bytes of device zeros; workspace: >= rtp_wgrad_s2d_workspace_bytes.
 * replaces: autograd's conv3d weight gradient for the fuse / transition convs (hr_util/hr3d.py:162-197, :286-331). */
int rtp_wgrad_s2d_supported(int32_t Cin, int32_t NP, int32_t Z, int32_t X, int32_t Y);
int64_t rtp_wgrad_s2d_workspace_bytes(int32_t NP, int32_t nsm);
int rtp_wgrad_s2d(rtp_p8 xs, rtp_p8 dy, int32_t Cin, int32_t NP, const void* zero_page, float* workspace,
                  int32_t* nsplit_out, void* stream);
int rtp_wgrad_s2d_reduce(const float* workspace, int32_t nsplit, int32_t Cin, int32_t NP, float* dW, int32_t co_n,
                         int32_t accumulate, void* stream);
int rtp_weight_s2d_expand(const float* w, float* w_s2d, int32_t Cout, int32_t Cin, void* stream);
int rtp_weight_s2d_fold(const float* dw_s2d, float* dw, int32_t Cout, int32_t Cin, int32_t accumulate, void* stream);
/* Sibling stride-2 fuse convs that read the same tensor (fuse_layers[i][0][0], i = 1..3, of one HighResolutionModule,
 * hr_util/hr3d.py:159-203: each is GroupNorm -> Conv3d(stride 2) on the full-resolution branch) share ONE space-to-depth
 * view of xhat = (x - mean) * rstd (rtp_gn_apply_s2d with gamma = 1, beta = 0); the GroupNorm affine of each sibling is folded
 * into its conv (csrc/s2d_shared.cu):
 *   rtp_s2d_fold_weights: w_s2d = expand(W * diag(gamma)) [Cout][8*Cin][27]; bias_cls[8][Cout] = the beta term per border
 *     class (bit 2: zo == 0, bit 1: xo == 0, bit 0: yo == 0 — the output positions whose k = 0 taps leave the volume);
 *   rtp_s2d_border_bias: r (P8, N = 1, broadcast over samples through n_stride = 0 as the conv's `res`) =
 *     bias_cls[class(pos)][c] - bias_cls[0][c]; bias_cls[0] goes in as the conv's per-channel bias;
 *   rtp_s2d_fold_wgrad: from dw_xhat = wgrad(xhat view, dy) and the box sums of dy (workspace >=
 *     rtp_s2d_box_sums_workspace_bytes(N, ceil(Cout/8))): dW (=|+=) gamma*dw_xhat + beta*T, dgamma (=|+=) sum W*dw_xhat,
 *     dbeta (=|+=) sum W*T, with T[co][tap] = sum of dy[co] over the positions whose tap is inside the volume.
 * The siblings' dgrads (gamma-folded weights) accumulate into one dL/dxhat view and ONE rtp_gn_bwd_*_s2d call with
 * gamma = 1 (dgamma = dbeta = NULL) produces dL/dx.  replaces: autograd through those GroupNorm + Conv3d pairs. */
int rtp_s2d_fold_weights(const float* w, const float* gamma, const float* beta, float* w_s2d, float* bias_cls, int32_t Cout,
                         int32_t Cin, void* stream);
int rtp_s2d_border_bias(const float* bias_cls, rtp_p8 r, int32_t C, void* stream);
int64_t rtp_s2d_box_sums_workspace_bytes(int32_t N, int32_t C8);
int rtp_s2d_fold_wgrad(rtp_p8 dy, const float* dw_xhat, const float* w, const float* gamma, const float* beta, float* workspace,
                       float* dW, float* dgamma, float* dbeta, int32_t Cout, int32_t Cin, int32_t accumulate_w,
                       int32_t accumulate_gb, void* stream);

/* Backward of the regression branch's last conv from the SPARSE loss gradient (csrc/head_sparse.cu).  CenterHead.loss gathers
 * the regression output at the target voxels only (center_head.py:244-270), so d_reg — written by rtp_head_loss — is zero
 * except at the <= M voxels ind[n][0..M) of each sample (reference flat index z*Y*X + y*X + x).  Computes, without touching the
 * rest of the volume: dt = [t_in > 0] * conv_transpose(d_reg, w) (dt is zero-filled first; Cin channels), dW (= or +=) and
 * db (= or +=) of the conv y = conv3x3x3(t_in, w) + b, w fp32 [R][Cin][3][3][3].  M <= 64, R <= 64.
 * replaces: autograd through SepHead's last Conv (center_head.py:95-104) for the `reg` branch. */
int64_t rtp_reg_head_bwd_sparse_workspace_bytes(int32_t N, int32_t M);
int rtp_reg_head_bwd_sparse(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R, int32_t Cin,
                            rtp_p8 dt, float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b, void* workspace,
                            void* stream);
/* The same for a dt whose first ceil(Cin / 8) chunks are ALREADY zero (rtp_zero_chunks issued earlier, e.g. on a side stream
 * beside the head's convolutions): skips the zero-fill, which otherwise sits at the serial start of the backward pass. */
int rtp_reg_head_bwd_sparse_prezeroed(rtp_p8 d_reg, rtp_p8 t_in, const int64_t* ind, int32_t M, const float* w, int32_t R, int32_t Cin,
                            rtp_p8 dt, float* dW, int32_t accumulate_w, float* db, int32_t accumulate_b, void* workspace,
                            void* stream);
/* Zero-fills all t.C8 chunk volumes of every sample of a P8 tensor / channel view (pad ring included; chunk volumes must be dense). */
int rtp_zero_chunks(rtp_p8 t, void* stream);

/* ---- branch exchange ---------------------------------------------------------------------------------------
 * replaces: the fuse sum of HighResolutionModule.forward (hr_util/hr3d.py:213-227) and the upsample+cat of
 * HRNet3D.forward (hrnet3d.py:37-42): out = [relu]( sum_i same[i] + sum_j trilinear_up(low[j]) + bias ),
 * F.interpolate(mode="trilinear", align_corners=True) semantics; fp32 accumulation, one bf16 store. */
typedef struct {
  rtp_p8 out;
  int32_t C;
  int32_t n_same, n_low;
  rtp_p8 same[4];
  rtp_p8 low[3];
  const float* bias; /* [C] or NULL */
  int32_t relu;
} rtp_fuse_desc;
int rtp_fuse_sum(const rtp_fuse_desc* d, void* stream);

/* The final concat + 1x1 conv of HRNet3D as ONE streaming GEMM (csrc/conat.cu):
 *   out = [relu]( W * cat(x0, up(low[0]), up(low[1]), up(low[2])) + bias ),   up = trilinear, align_corners=True
 * replaces: `x = torch.cat([x0, F.interpolate(x1..x3, size, mode='trilinear', align_corners=True)], 1); final_conv(x)`
 * (det3d/models/backbones/hrnet3d.py:37-42).  The chunks of x0 reach shared memory by bulk copies, the chunks of the
 * upsampled branches are interpolated into the UMMA A operand by the kernel; the concat is never materialised.
 * w: rtp_weight_pack(mode 0) of the whole [Cout][K] weight, K = c_x0 + sum c_low (each a multiple of 16, cat order).
 * Supported when rtp_conat_supported() returns 1 (Cout <= 128, weights + two operand stages + blend rows fit shared
 * memory, padded row length Y + 2 >= 43); callers fall back to per-branch rtp_conv_pw + rtp_fuse_sum otherwise. */
typedef struct {
  rtp_p8 x0, out;
  rtp_p8 low[3];
  int32_t n_low;      /* 1..3 */
  int32_t c_x0;       /* channels taken from x0 */
  int32_t c_low[3];   /* channels taken from each low term */
  const void* w;      /* packed weights [K/8][NP][8] bf16 */
  const float* bias;  /* [NP] fp32 or NULL */
  int32_t K, NP, out_c8, relu;
} rtp_conat_desc;
int rtp_conat_supported(const rtp_conat_desc* d);
int rtp_conat_fwd(const rtp_conat_desc* d, void* stream);
/* dlow (=|+=) transpose-of-trilinear-upsample applied to dout, as three separable 1-D gather passes (y, x, z;
 * deterministic); workspace >= rtp_upsample_bwd_workspace_bytes(dout, dlow, C), 16-byte aligned */
int64_t rtp_upsample_bwd_workspace_bytes(rtp_p8 dout, rtp_p8 dlow, int32_t C);
int rtp_upsample_bwd(rtp_p8 dout, rtp_p8 dlow, int32_t C, int32_t accumulate, void* workspace, void* stream);
/* dst (=|+=) src [* (mask > 0)]   — gradient pass-through of the fuse sum / residual add */
int rtp_grad_add(rtp_p8 src, rtp_p8 mask, rtp_p8 dst, int32_t C, int32_t accumulate, void* stream);
/* out[c] = sum over (n, voxels) of x[n][c] (fp32) — bias gradients */
int rtp_channel_sum(rtp_p8 x, int32_t C, float* out, int32_t accumulate, float* workspace, void* stream);

/* ---- 1 -> C stem (conv1 of layer1 when Cin != Cout; hr_util/common.py:113-116,140) ------------------------
 * y[c] = w[c] * x + b[c] for single-channel input (x is channel 0 of a P8 tensor). */
int rtp_stem_fwd(rtp_p8 x, const float* w, const float* b, int32_t C, rtp_p8 y, void* stream);
int rtp_stem_bwd(rtp_p8 x, rtp_p8 dy, int32_t C, float* dw, float* db, int32_t accumulate, float* workspace,
                 void* stream);

/* ---- CenterHead loss -----------------------------------------------------------------------------------------
 * replaces: CenterHead.loss (center_head.py:244-270) -> FastFocalLoss.forward (losses/centernet_loss.py:34-54),
 * RegLoss.forward (:17-24), _transpose_and_gather_feat (core/utils/center_utils.py:113-117).
 * hm, reg: raw head outputs (P8).  tgt_hm fp32 [N][ncls][Z][Y][X]; ind int64 [N][M] (flat z*Y*X+y*X+x);
 * mask uint8 [N][M]; cat int64 [N][M]; anno fp32 [N][M][R].
 * out (fp32, device): [0]=loss [1]=hm_loss [2]=loc_loss [3]=num_pos [4..4+R)=loc_loss_elem.
 * d_hm / d_reg (P8, may be NULL): gradients of `loss * grad_scale` w.r.t. the raw outputs. */
int64_t rtp_head_loss_workspace_bytes(int32_t N, int32_t ncls, int32_t Z, int32_t Y, int32_t X);
int rtp_head_loss(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* tgt_hm, const int64_t* ind,
                  const uint8_t* mask, const int64_t* cat, const float* anno, int32_t M, float weight,
                  const float* code_weights, float grad_scale, float* out, rtp_p8 d_hm, rtp_p8 d_reg,
                  void* workspace, void* stream);
/* The same with options.  Chunks of d_hm behind the ceil(ncls / 8) class chunks (d_hm.C8 larger than that: a K-padding chunk for
 * the tensor-core dgrad that follows) are zero-filled in the same pass.  flags & RTP_LOSS_SPARSE_DREG: d_reg is written (cleared,
 * then accumulated) at the N * M target voxels ind[n][m] only and left untouched everywhere else — for a consumer that reads it
 * there only (rtp_reg_head_bwd_sparse); saves the dense zero-fill of the R-channel gradient. */
#define RTP_LOSS_SPARSE_DREG 1
int rtp_head_loss_flags(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* tgt_hm, const int64_t* ind,
                        const uint8_t* mask, const int64_t* cat, const float* anno, int32_t M, float weight,
                        const float* code_weights, float grad_scale, float* out, rtp_p8 d_hm, rtp_p8 d_reg,
                        int32_t flags, void* workspace, void* stream);

/* ---- keypoint decode -------------------------------------------------------------------------------------------
 * replaces: CenterHead.predict + post_processing (center_head.py:272-360): per (sample, class) arg-max of
 * sigmoid(hm) over Z*Y*X (lowest reference flat index wins ties), gather of the regression row, metric
 * coordinates (idx + reg) * voxel + range.
 * out_index int32 [N][ncls]; out_score fp32 [N][ncls]; out_xyz fp32 [N][ncls][R] (R = 3 or 45).
 * voxel_xyz / range_xyz are HOST pointers to 3 floats each (configuration, read at call time). */
int rtp_decode(rtp_p8 hm, rtp_p8 reg, int32_t ncls, int32_t R, const float* voxel_xyz, const float* range_xyz,
               int32_t* out_index, float* out_score, float* out_xyz, void* stream);

/* ---- deformable convolution v1 (2-D) ------------------------------------------------------------------------
 * replaces: deform_conv_forward_cuda / deform_conv_backward_input_cuda / deform_conv_backward_parameters_cuda
 * (det3d/ops/dcn/src/deform_conv_cuda.cpp:152-488; kernels deform_conv_cuda_kernel.cu:190-465; bound at
 * deform_conv_cuda.cpp:687-701 and called from det3d/ops/dcn/deform_conv.py:52,77,87).
 * fp32 NCHW tensors as in the reference op; offset channel order [dg][kh*kw][dy,dx]; groups = 1. */
int rtp_dcn_fwd(const float* x, const float* offset, const float* w, float* y, int32_t N, int32_t C, int32_t H,
                int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil,
                int32_t dg, void* stream);
int rtp_dcn_bwd_input(const float* x, const float* offset, const float* w, const float* dy, float* dx, float* doffset,
                      int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw,
                      int32_t stride, int32_t pad, int32_t dil, int32_t dg, void* stream);
int rtp_dcn_bwd_weight(const float* x, const float* offset, const float* dy, float* dw, int32_t N, int32_t C,
                       int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                       int32_t dil, int32_t dg, float scale, void* stream);

/* Tensor-core forward, step 1 of 2: deformed im2col as a bf16 P8 volume whose z axis is the tap index,
 * dst[n][c/8][t = i*kw+j][wo][ho][c%8] (dst: N, C8*8 >= C, Z = kh*kw, Y = Ho, X = Wo; halo left untouched, i.e. zero).
 * mask = NULL for v1.  Step 2 is rtp_conv with taps {(tz = t, tx = 0, ty = 0, wt = t)}, row grid (1, Wo, Ho) and the weight
 * packed by rtp_weight_pack(ntaps = kh*kw): the contraction over C*kh*kw runs on tcgen05 with fp32 accumulation instead of
 * the reference's fp32 `columns` x cuBLAS GEMM (deform_conv_cuda.cpp:196-247).  (C / dg) % 8 == 0, N*dg*kh*kw <= 65535. */
int rtp_dcn_sample_p8(const float* x, const float* offset, const float* mask, rtp_p8 dst, int32_t N, int32_t C, int32_t H,
                      int32_t W, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t dil, int32_t dg, void* stream);

/* Tensor-core backward, last step.  ds: bf16 P8 gradient of the sample volume (layout of rtp_dcn_sample_p8; produced by
 * kh*kw single-tap rtp_conv launches from dy with the dgrad-packed weight, output plane oz0 = t).  Overwrites dx
 * (zero + atomic scatter), doffset and, when mask != NULL, dmask.  The weight gradient of this path is rtp_wgrad over
 * (sample volume, dy) with the tap list {(t, 0, 0)}; rtp_dcn_bias_grad ACCUMULATES scale * sum_{n,pix} dy into dbias. */
int rtp_dcn_col2im_p8(const float* x, const float* offset, const float* mask, rtp_p8 ds, float* dx, float* doffset, float* dmask,
                      int32_t N, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                      int32_t dil, int32_t dg, void* stream);
int rtp_dcn_bias_grad(const float* dy, float* dbias, int32_t N, int32_t Cout, int32_t npix, float scale, void* stream);

/* ---- deformable convolution v2 ("modulated", 2-D) ---------------------------------------------------------------
 * replaces: modulated_deform_conv_cuda_forward / modulated_deform_conv_cuda_backward
 * (det3d/ops/dcn/src/deform_conv_cuda.cpp:490-684; kernels deform_conv_cuda_kernel.cu:467-867; bound at
 * deform_conv_cuda.cpp:687-701 and called from det3d/ops/dcn/deform_conv.py:141,161).
 * As v1 plus mask fp32 [N][dg*kh*kw][Ho][Wo] multiplying every sample and an optional bias[Cout] (NULL = none).
 * bwd_input overwrites dx / doffset / dmask; bwd_weight ACCUMULATES scale * gradient into dw and (if not NULL) dbias,
 * like the reference, whose autograd function passes zero-filled tensors (deform_conv.py:156-160). */
int rtp_mdcn_fwd(const float* x, const float* offset, const float* mask, const float* w, const float* bias, float* y,
                 int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride,
                 int32_t pad, int32_t dil, int32_t dg, void* stream);
int rtp_mdcn_bwd_input(const float* x, const float* offset, const float* mask, const float* w, const float* dy, float* dx,
                       float* doffset, float* dmask, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh,
                       int32_t kw, int32_t stride, int32_t pad, int32_t dil, int32_t dg, void* stream);
int rtp_mdcn_bwd_weight(const float* x, const float* offset, const float* mask, const float* dy, float* dw, float* dbias,
                        int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cout, int32_t kh, int32_t kw, int32_t stride,
                        int32_t pad, int32_t dil, int32_t dg, float scale, void* stream);

/* ---- "next" rows around the path --------------------------------------------------------------
 * N2 fused optimizer step on flat fp32 buffers.  replaces: OptimizerHook.clip_grads (clip_grad_norm_, max_norm 35;
 * det3d/torchie/trainer/hooks/optimizer.py:9-24) + OptimWrapper.step (decoupled weight decay p *= 1 - wd*lr on every
 * parameter, bn_wd=True; det3d/solver/fastai_optim.py:158-174) + torch.optim.Adam(betas=(mom, 0.99), eps=1e-8).step().
 * step >= 1 is the Adam time step; max_norm <= 0 disables clipping; workspace >= rtp_adam_workspace_bytes();
 * grad_norm_out (optional, device) receives the global L2 norm before clipping. */
int64_t rtp_adam_workspace_bytes(void);
int rtp_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float wd, int32_t step, float max_norm, float* workspace, float* grad_norm_out, void* stream);
/* Same step with the hyper-parameters read from DEVICE memory: hyper = {lr, beta1, beta2, eps, wd, 1 - beta1^t,
 * sqrt(1 - beta2^t), max_norm} (8 floats).  Lets a CUDA graph that captured the whole training step be replayed along
 * the one-cycle schedule (learning_schedules_fastai.py:53-95) by rewriting 32 bytes between replays. */
int rtp_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, const float* hyper, float* workspace,
                      float* grad_norm_out, void* stream);
/* N1 CenterNet target assignment on the device.  replaces: AssignLabelPose / AssignLabelPose2.__call__
 * (det3d/datasets/pipelines/pose.py:186-255, :385-452) with gaussian3D / draw_gaussian3D
 * (det3d/core/utils/center_utils.py:67-91).  poses: device fp64 [B][15][3] metres; voxel_xyz (fp64) and range_xyz (fp32)
 * are HOST pointers to 3 values.  Outputs (device): hm fp32 [B][ncls][Z][Y][X], ind int64 [B][M], mask uint8 [B][M],
 * cat int64 [B][M], anno fp32 [B][M][R] with (ncls, M, R) = (1, 1, 45) when one_hm else (15, 15, 3). */
int rtp_assign_targets(const double* poses, int32_t B, int32_t Z, int32_t Y, int32_t X, int32_t one_hm, int32_t radius,
                       const double* voxel_xyz, const float* range_xyz, float* hm, int64_t* ind, uint8_t* mask,
                       int64_t* cat, float* anno, void* stream);

/* N4 pose-error metrics on the device.  replaces: PJPE / ABS_PJPE (eval_util.py:5-11) and the per-sequence, per-joint
 * means of CRUW_POSE_Dataset.evaluation (det3d/datasets/cruw_pose/cruw_pose.py:277-295).
 * pred_xyz: device fp32 [N][J][3] (rtp_decode's out_xyz: J = 15 rows of 3 for `hr3d`, one row of 45 for `one_hm`);
 * gt_xyz: device fp64 [N][J][3] (the label file's floats).  out_rel = root-relative error (joint 0 subtracted on both
 * sides), out_abs = absolute error, both fp64 [N][J] in metres, computed in the reference's fp64 operation order.
 * rtp_pjpe_seq_mean: seq_index int32 [N] in [0,S) -> means over each sequence's frames x 1000 (millimetres),
 * fp64 [S][J] each, and the frame count per sequence (0 frames -> 0). */
int rtp_pjpe(const float* pred_xyz, const double* gt_xyz, int32_t N, int32_t J, double* out_rel, double* out_abs, void* stream);
int rtp_pjpe_seq_mean(const double* rel, const double* abs_, const int32_t* seq_index, int32_t N, int32_t J, int32_t S,
                      double* out_rel_mm, double* out_abs_mm, int32_t* out_count, void* stream);

/* ---- flat-buffer helpers for the data-parallel step ----------------------------------------------------------
 * replaces: _allreduce_coalesced's flatten / div_ / copy-back (det3d/core/utils/dist_utils.py:8-28); the
 * collective itself is ncclAllReduce issued by torch.distributed on the same flat buffer. */
int rtp_scale_f32(float* buf, int64_t n, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTPOSE_B200_H_ */
