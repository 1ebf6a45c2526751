// conv_k3s1.cu — plane-streaming 3x3x3 stride-1 pad-1 conv3d (forward, and dgrad with flipped weights) on tcgen05.
//
// The dominant shape of HRRadarPose (8 full-resolution 32->32 convs = 39-63 % of the forward FLOPs, plus their
// dgrads).  Design, following the measurements in profiles/r01_umma_probe.txt:
//
//  * UMMA reads its SMEM operands at 128 B/clk/SM, so an N=32 GEMM is capped at ~35 % of the tensor peak.  The three
//    z-taps are therefore stacked into GEMM N:  B = W[(ky,kx,ci), (kz,co)]  (N = 3*Cout = 96 -> 86 % bound), and the
//    product of input plane z lands in the TMEM accumulator blocks of output planes z-1, z, z+1, which are laid
//    out contiguously (block b = 32 columns of plane zo0+b; 16 planes x 32 columns = all 512 TMEM columns).
//  * In the P8 layout an in-plane tap (ky,kx) is a linear shift, so the A operand of every tap is the SAME
//    shared-memory stage addressed with a start offset of (kx*Yp + ky)*16 B (SWIZZLE_NONE K-major descriptor):
//    each input plane is fetched ONCE per tile by 1-D bulk async copies (no im2col, no tensor map, zero padding
//    comes from the layout's zero ring).
//  * Persistent CTAs (one per SM) walk (sample, 128-position tile, z-chunk) units.  Warp 0 = bulk-copy producer,
//    warp 1 = MMA issuer, warps 2-9 = two epilogue groups that drain finished planes (tcgen05.ld -> bias /
//    residual / ReLU / mask / accumulate -> bf16 stores) while the MMAs of later planes and of the next unit run.
//  * Wide K (the 128-channel head) is processed in passes of KG channels with double-buffered weight slices; the
//    accumulators stay in TMEM across passes.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * N*Z*Y*X * Cout * Cin * 27.
#include <cstdlib>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int kThreads = 320;   // LANES = 1: warp 0 producer, warp 1 MMA, warps 2-5 / 6-9 epilogue groups 0 / 1
constexpr int kThreads2 = 384;  // LANES = 2: warps 0-1 producers, 2-3 MMA issuers, 4-7 / 8-11 the epilogue group of lane 0 / 1
// Register cap of the dual-lane instantiations, expressed as a launch bound.  -DRTP_K3S1_DUAL_BOUND=512 compiles them to
// <= 128 registers (no spills: 48 k of the SM's 64 k registers, room for an element-wise CTA of another stream next to the
// persistent conv CTA); measured, it changes nothing in the step (20.99 vs 21.03 ms) and costs the kernel 1 %, so the
// default stays at the launch size (145-149 registers).
#ifndef RTP_K3S1_DUAL_BOUND
#define RTP_K3S1_DUAL_BOUND 384
#endif
constexpr int kMaxBlocks = 32;
constexpr int kMaxStages = 8;
#ifdef RTP_K3S1_DEBUG
constexpr bool kDbg = true;   // cycle counters of the MMA warp (tools/dbg_k3s1.py builds with -DRTP_K3S1_DEBUG)
#else
constexpr bool kDbg = false;  // the counters cost a CS2R per step boundary: compiled out of the production kernel
#endif

struct K3 {
  P8 in, out, res, mask;
  const bf16* w;
  const float* bias;
  int NPo, N3, out_c8, relu, accumulate, has_res, has_mask;
  int KG, npass;   // channels per pass, passes (KG * npass == K)
  int PW;          // stage width in positions: 128 + 2*Yp + 2
  int ntile;       // 128-position tiles per plane
  int ZC, nzc;     // output planes per z-chunk, z-chunks
  int nunits;
  // LANES = 1, one z-chunk: the units of the last, partly filled round (index >= split_from) are issued as two half-depth units
  // each (output planes [0, split_zh) and [split_zh, Z)), so that round costs (Z/2 + 1) / Z of a full one
  int split_from, split_zh;
  const int* unit_list;   // optional: the unit ids to process (rtp_active_units) and their count, else all nunits
  const int* unit_count;
  int nstages;
  uint32_t stage_bytes, wbuf_bytes, wtap_bytes, wtap_stride;  // per-pass weight slice: 9 copies of wtap_bytes
  long long* dbg;  // optional [grid][8] cycle counters (RTP_K3S1_DEBUG): MMA-warp wait/issue breakdown
  // fused per-(sample, channel) statistics of the stored result (STAT template argument, out_c8 <= 4):
  //   1: sum v, sum v*v        (the next GroupNorm's mean / variance)
  //   2: sum v, sum v*aux      (GroupNorm backward: v = dL/d(normalised x), aux = x)
  P8 stat_aux;
  float* stat_ws;  // [grid*8 warp slabs][N][64] then [grid CTA slabs][N][64]
  uint32_t pre_off;      // LANES = 2: byte offset (dynamic smem) of the per-thread residual / aux prefetch ring, 0 = off
  uint16_t tapmask[32];  // per K pass: bit t9 set = in-plane tap t9 has non-zero weights (structurally sparse weights)
};

// LANES = 2 ("dual issue"): the issuing thread, not the tensor pipe, bounds the single-lane kernel — per plane-step it
// spends ~875 cycles issuing 18 MMAs and ~560 in barrier waits / commits (tools/dbg_k3s1.py) while the pipe needs 1008 and
// idles whenever its short queue drains.  With two lanes the CTA runs TWO independent producer / issuer / epilogue chains
// that share the weights and the tensor pipe: lane L owns z-chunk L of the unit (output planes [L*ZC, (L+1)*ZC), its own
// half of TMEM, its own half of the stage ring), so while one issuer waits or commits the other keeps the pipe fed.  Every
// output is still accumulated by ONE issuer in a fixed order (results stay run-to-run identical); the price is the halo
// plane at the chunk boundary (18 instead of 16 plane loads, +8 % MMA cycles at Z = 16).  Single K pass only.
template <int KS, int STAT, int LANES>  // KS = KG / 16: k16 steps per tap (1 or 2); STAT: fused statistics mode (0 = off)
__global__ void __launch_bounds__(LANES == 2 ? RTP_K3S1_DUAL_BOUND : kThreads, 1) conv_k3s1_kernel(const __grid_constant__ K3 p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_full_s[kMaxStages], bar_empty_s[kMaxStages], bar_wfull[2], bar_wempty[2];
  __shared__ uint64_t bar_acc_full_s[kMaxBlocks], bar_acc_empty_s[kMaxBlocks];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // role of this warp and the lane (z-chunk chain) it serves
  const int role = LANES == 2 ? (warp < 2 ? 0 : (warp < 4 ? 1 : 2)) : (warp == 0 ? 0 : (warp == 1 ? 1 : 2));
  const int L = LANES == 2 ? (role == 2 ? (warp - 4) >> 2 : (warp & 1)) : 0;
  const int S = p.nstages / LANES;  // stages per lane
  const int nwbuf = p.npass > 1 ? 2 : 1;
  uint8_t* wbuf = smem;
  uint8_t* stages = smem + (size_t)nwbuf * p.wbuf_bytes + (size_t)L * S * p.stage_bytes;
  uint64_t* bar_full = bar_full_s + L * S;
  uint64_t* bar_empty = bar_empty_s + L * S;
  uint64_t* bar_acc_full = bar_acc_full_s + L * (kMaxBlocks / 2);
  uint64_t* bar_acc_empty = bar_acc_empty_s + L * (kMaxBlocks / 2);
  const int Yp = p.in.Yp, Z = p.in.Z;
  const int kch = p.KG >> 3;

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&bar_full_s[s], 1); mbar_init(&bar_empty_s[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_wfull[s], 1); mbar_init(&bar_wempty[s], 1); }
    for (int b = 0; b < kMaxBlocks; ++b) { mbar_init(&bar_acc_full_s[b], 1); mbar_init(&bar_acc_empty_s[b], 128); }
    mbar_fence_init();
  }
  if (warp == (LANES == 2 ? 2 : 1)) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t)L * 256u;  // lane L accumulates in its own 256 columns

  // returns false when this lane has nothing to do in unit u (LANES = 2, split last round: one lane per CTA)
  // with a unit list, entry b = n * ntile + tile stands for all z-chunks of that (sample, tile)
  const int zmul = (LANES == 1 && p.unit_list) ? p.nzc : 1;
  const int nloop = p.unit_list ? __ldg(p.unit_count) * zmul : p.nunits;
  auto unit_of = [&](int ku) {
    if (!p.unit_list) return ku;
    if (zmul == 1) return p.unit_list[ku];
    const int b = p.unit_list[ku / zmul], zc = ku - (ku / zmul) * zmul;
    return ((b / p.ntile) * p.nzc + zc) * p.ntile + b % p.ntile;
  };
  auto decode = [&](int u, int& n, int& zo0, int& zo1, int& tile) -> bool {
    if constexpr (LANES == 2) {  // a unit is (sample, tile); the z-chunk is the lane
      int uu = u;
      bool active = true;
      if (u >= p.split_from) {  // last round: unit uu is shared by two CTAs, each running ONE lane with the tensor pipe to itself
        const int h = u - p.split_from;
        uu = p.split_from + (h >> 1);
        active = (h & 1) == L;
      }
      tile = uu % p.ntile;
      n = uu / p.ntile;
      zo0 = L * p.ZC;
      zo1 = min(Z, zo0 + p.ZC);
      return active;
    } else if (u < p.split_from) {
      tile = u % p.ntile;
      const int r = u / p.ntile;
      n = r / p.nzc;
      zo0 = (r % p.nzc) * p.ZC;
      zo1 = min(Z, zo0 + p.ZC);
    } else {  // half-depth units of the last round (nzc == 1)
      const int h = u - p.split_from, uu = p.split_from + (h >> 1);
      tile = uu % p.ntile;
      n = uu / p.ntile;
      zo0 = (h & 1) ? p.split_zh : 0;
      zo1 = (h & 1) ? Z : p.split_zh;
    }
    return true;
  };

  if (role == 0) {
    // ============================================================ producer
    // lane 0 runs the barrier protocol; the copies of a step are issued by several lanes at once (one thread issuing
    // bulk copies back to back is limited to ~8 GB/s per SM, profiles/r01_bulk_probe.txt)
    {
      uint32_t it = 0, wit = 0;
      bool w_loaded = false;
      for (int ku = blockIdx.x; ku < nloop; ku += gridDim.x) {
        const int u = unit_of(ku);
        int n, zo0, zo1, tile;
        if (!decode(u, n, zo0, zo1, tile)) continue;
        const int iz0 = max(0, zo0 - 1), iz1 = min(Z, zo1 + 1);
        const int64_t qoff = ((int64_t)tile * 128 - 1) * 8;  // first staged position = q0 - Yp - 1, q0 = Yp + tile*128
        const bf16* in_n = p.in.ptr + (int64_t)n * p.in.n_stride + qoff;
        for (int g = 0; g < p.npass; ++g) {
          if ((p.npass > 1 || !w_loaded) && L == 0) {
            const int wb = wit & 1;
            if (lane == 0) {
              mbar_wait(&bar_wempty[wb], ((wit >> 1) & 1) ^ 1);
              mbar_arrive_expect_tx(&bar_wfull[wb], p.wbuf_bytes);
            }
            __syncwarp();
            if (lane < 9)
              bulk_g2s(wbuf + (size_t)wb * p.wbuf_bytes + (size_t)lane * p.wtap_bytes,
                       reinterpret_cast<const uint8_t*>(p.w) + (size_t)lane * p.wtap_stride + (size_t)g * p.wtap_bytes,
                       p.wtap_bytes, &bar_wfull[wb]);
            ++wit;
            w_loaded = true;
          }
          for (int iz = iz0; iz < iz1; ++iz) {
            const int s = it % S;
            if (lane == 0) {
              mbar_wait(&bar_empty[s], ((it / S) & 1) ^ 1);
              mbar_arrive_expect_tx(&bar_full[s], p.stage_bytes);
            }
            __syncwarp();
            uint8_t* dst = stages + (size_t)s * p.stage_bytes;
            if (lane < kch)
              bulk_g2s(dst + (size_t)lane * p.PW * 16,
                       in_n + (int64_t)(g * kch + lane) * p.in.c_stride + (int64_t)iz * p.in.plane_elems(), p.PW * 16,
                       &bar_full[s]);
            ++it;
          }
        }
      }
    }
  } else if (role == 1) {
    // ============================================================ MMA issuer
    // Loop control, barrier waits and descriptor arithmetic run warp-uniformly (uniform datapath, no per-MMA
    // reconvergence); only the tcgen05.mma / commit instructions are predicated on the leader lane.
    {
      const bool leader = lane == 0;
      uint32_t it = 0, wit = 0, fresh_mask = 0;  // bit b: parity of the number of first-writes to block b so far
      long long t_empty = 0, t_full = 0, t_issue = 0, t_first = 0, t_commit = 0, t0 = kDbg ? clock64() : 0, tk = 0;
      bool w_ready = false;
      uint32_t wcur = 0;
      const uint32_t idesc1 = idesc_bf16(128, p.NPo, 0, 0), idesc2 = idesc_bf16(128, 2 * p.NPo, 0, 0),
                     idesc3 = idesc_bf16(128, 3 * p.NPo, 0, 0);
      // smem descriptor halves (SWIZZLE_NONE, K-major): lo = start>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version bit
      const uint32_t a_lo_c = (uint32_t)p.PW << 16, b_lo_c = (uint32_t)p.N3 << 16;
      const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;
      const uint32_t a_k16 = 2u * p.PW, b_k16 = 2u * p.N3, b_tap16 = p.wtap_bytes >> 4;
      const uint32_t stage0 = smem_u32(stages), wbase0 = smem_u32(wbuf);
      auto mk_desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
      for (int ku = blockIdx.x; ku < nloop; ku += gridDim.x) {
        const int u = unit_of(ku);
        int n, zo0, zo1, tile;
        if (!decode(u, n, zo0, zo1, tile)) continue;
        const int iz0 = max(0, zo0 - 1), iz1 = min(Z, zo1 + 1);
        for (int g = 0; g < p.npass; ++g) {
          if (p.npass > 1 || !w_ready) {
            wcur = wit & 1;
            mbar_wait(&bar_wfull[wcur], (wit >> 1) & 1);
            w_ready = true;
          }
          for (int iz = iz0; iz < iz1; ++iz) {
            const int s = it % S;
            const int lo = max(iz - 1, zo0), hi = min(iz + 1, zo1 - 1);
            if constexpr (kDbg) tk = clock64();
            if (g == 0) {  // blocks written for the first time in this unit must have been drained by the epilogue
              if (iz + 1 <= hi) {
                const int b = iz + 1 - zo0;
                mbar_wait(&bar_acc_empty[b], (fresh_mask >> b) & 1);
                fresh_mask ^= 1u << b;
              }
              if (iz == 0) {
                mbar_wait(&bar_acc_empty[0], fresh_mask & 1);
                fresh_mask ^= 1u;
              }
            }
            if constexpr (kDbg) { const long long t1 = clock64(); t_empty += t1 - tk; tk = t1; }
            mbar_wait(&bar_full[s], (it / S) & 1);
            fence_after_sync();
            if constexpr (kDbg) { const long long t1 = clock64(); t_full += t1 - tk; tk = t1; }
            // Descriptors are built from precomputed halves: only the 14-bit start-address field (16-byte units) of the
            // low word changes between MMAs, so one elected thread sustains the issue rate (~10 instructions / MMA).
            const uint32_t dcol = tmem + (uint32_t)(lo - zo0) * p.NPo;
            const uint32_t boff16 = (uint32_t)(lo - (iz - 1)) * p.NPo;  // N-slice offset in 16-byte rows
            const int nblk = hi - lo + 1;
            const uint32_t idesc = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
            const uint32_t a_lo = a_lo_c + ((stage0 + (uint32_t)s * p.stage_bytes) >> 4);
            const uint32_t b_lo = b_lo_c + ((wbase0 + wcur * p.wbuf_bytes) >> 4);
            // every accumulator block is handed over zeroed by the epilogue (tcgen05.st after the drain, and once at
            // kernel start), so all MMAs accumulate: no overwrite / split first tap, one uniform N = 96 stream
            if constexpr (kDbg) { const long long t1 = clock64(); t_first += t1 - tk; tk = t1; }
            const uint32_t tm = p.tapmask[g];
            if (elect_one()) {  // elect.sync: the compiler knows exactly one lane runs this block (no per-MMA waterfall)
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) {
                if (!((tm >> t9) & 1u)) continue;  // all-zero tap of this pass (space-to-depth weights)
                const uint32_t at = a_lo + (uint32_t)((t9 / 3) * Yp + (t9 % 3));
                const uint32_t bt = b_lo + boff16 + t9 * b_tap16;
#pragma unroll
                for (int k16 = 0; k16 < KS; ++k16) {
                  mma_ss(dcol, mk_desc(at + k16 * a_k16, a_hi), mk_desc(bt + k16 * b_k16, b_hi), idesc, 1u);
                }
              }
              if constexpr (kDbg) { const long long t1 = clock64(); t_issue += t1 - tk; tk = t1; }
              mma_commit(&bar_empty[s]);
              if (g == p.npass - 1) {
                if (iz - 1 >= zo0) mma_commit(&bar_acc_full[iz - 1 - zo0]);
                if (iz == Z - 1 && iz < zo1) mma_commit(&bar_acc_full[iz - zo0]);
              }
            }
            __syncwarp();
            if constexpr (kDbg) t_commit += clock64() - tk;
            ++it;
          }
          if (p.npass > 1) {
            if (leader) mma_commit(&bar_wempty[wcur]);
            ++wit;
          }
        }
      }
      if (kDbg && p.dbg && leader && L == 0) {
        long long* d = p.dbg + (size_t)blockIdx.x * 8;
        d[0] = clock64() - t0; d[1] = t_empty; d[2] = t_full; d[3] = t_issue; d[4] = it; d[5] = t_first; d[6] = t_commit;
      }
    }
  } else {
    // ============================================================ epilogue groups
    // LANES = 1: two groups take the even / odd accumulator blocks; LANES = 2: the group drains every block of its lane
    const int eg = LANES == 2 ? 0 : (warp - 2) >> 2;
    constexpr int kEgStep = LANES == 2 ? 1 : 2;
    const int ewarp = LANES == 2 ? warp - 4 : warp - 2;  // 0..7: index of this epilogue warp in the CTA
    const int lane_q = warp & 3;                    // TMEM lane quarter this warp may access
    const int r = lane_q * 32 + lane;               // GEMM row
    const uint32_t trow = tmem + ((uint32_t)(lane_q * 32) << 16);
    uint32_t full_mask = 0;  // bit b: parity of the number of times block b has been drained so far
    // hand every accumulator block of this group to the MMA warp zeroed (the MMAs only ever accumulate)
    for (int b = eg; b < p.ZC; b += kEgStep) {
      for (int c = 0; c < p.NPo; c += 16) tmem_st16_zero(trow + b * p.NPo + c);
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(&bar_acc_empty[b]);
    }
    // STAT: per-thread partial sums over the rows this thread stores; folded over the warp at the end of every unit
    // (fixed butterfly => deterministic) and added to this warp's private [N][64] slab in global memory
    float st0[STAT ? 32 : 1], st1[STAT ? 32 : 1];
    float* wslab = nullptr;
    if constexpr (STAT != 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) st0[i] = st1[i] = 0.f;
      wslab = p.stat_ws + ((size_t)blockIdx.x * 8 + ewarp) * p.out.N * 64;
      for (int n = 0; n < p.out.N; ++n) reinterpret_cast<float2*>(wslab + n * 64)[lane] = make_float2(0.f, 0.f);
    }
    // LANES = 2, opt-in (RTP_PRE_RING=1, see rtp_conv_k3s1): the residual (forward) / GroupNorm-input (dgrad statistics)
    // vectors of the NEXT block are fetched with cp.async into a two-deep per-thread ring in shared memory while the current
    // block is drained — an experiment against the exposed HBM latency of the register prefetch (ncu: 170 us with a residual,
    // 139 us without).  A thread only reads back what it copied itself, so cp.async.wait_group is the only synchronisation.
    constexpr bool kRing = LANES == 2 && STAT != 0;
    uint4* pre_ring = nullptr;
    bool use_ring = false;
    if constexpr (kRing) {
      use_ring = p.pre_off != 0 && (STAT == 2 || p.has_res);
      pre_ring = reinterpret_cast<uint4*>(smem + p.pre_off) + (size_t)(L * 128 + r) * 8;  // [2 buffers][4 chunks]
    }
    for (int ku = blockIdx.x; ku < nloop; ku += gridDim.x) {
        const int u = unit_of(ku);
      int n, zo0, zo1, tile;
      if (!decode(u, n, zo0, zo1, tile)) continue;
      const int q = Yp + tile * 128 + r;            // in-plane linear position (padded coordinates)
      const int xp = q / Yp, yp = q - xp * Yp;
      const bool ok = xp >= 1 && xp <= p.out.X && yp >= 1 && yp <= p.out.Y;
      const int64_t pos = (int64_t)q * 8;
      uint32_t rb = 0;  // ring buffer holding the CURRENT block's vectors
      auto ring_issue = [&](int oz2, uint32_t buf) {
        if constexpr (kRing) {
          if (ok) {
            const bf16* src = STAT == 2 ? p.stat_aux.ptr + (int64_t)n * p.stat_aux.n_stride + (int64_t)oz2 * p.stat_aux.plane_elems() + pos
                                        : p.res.ptr + (int64_t)n * p.res.n_stride + (int64_t)oz2 * p.res.plane_elems() + pos;
            const int64_t cs = STAT == 2 ? p.stat_aux.c_stride : p.res.c_stride;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < p.out_c8) cp_async16(pre_ring + buf * 4 + c, src + c * cs, true);
          }
          cp_async_commit();
        }
      };
      if (use_ring && zo0 + eg < zo1) ring_issue(zo0 + eg, 0);
      for (int oz = zo0 + eg; oz < zo1; oz += kEgStep) {
        const int b = oz - zo0;
        const int64_t plane = (int64_t)oz * p.out.plane_elems() + pos;
        bf16* out_row = p.out.ptr + (int64_t)n * p.out.n_stride + plane;
        const bf16* res_row = p.has_res ? p.res.ptr + (int64_t)n * p.res.n_stride + (int64_t)oz * p.res.plane_elems() + pos : nullptr;
        const bf16* mask_row = p.has_mask ? p.mask.ptr + (int64_t)n * p.mask.n_stride + (int64_t)oz * p.mask.plane_elems() + pos : nullptr;
        // issue the residual / mask / accumulate loads of the first four chunks BEFORE waiting for the accumulator,
        // so their latency overlaps the MMAs of this plane
        // (the STAT variants drop the operands their call sites never use, to keep the epilogue free of spills:
        //  1 = forward conv + residual, 2 = plain dgrad)
        constexpr bool kRes = STAT != 2, kMaskAcc = STAT == 0;
        uint4 pre_res[kRes ? 4 : 1], pre_mask[kMaskAcc ? 4 : 1], pre_acc[kMaskAcc ? 4 : 1], pre_aux[STAT == 2 ? 4 : 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if constexpr (kRes) pre_res[c] = make_uint4(0, 0, 0, 0);
          if constexpr (kMaskAcc) pre_mask[c] = pre_acc[c] = make_uint4(0, 0, 0, 0);
          if constexpr (STAT == 2) {
            pre_aux[c] = make_uint4(0, 0, 0, 0);
            if (ok && c < p.out_c8 && !use_ring)
              pre_aux[c] = ldg16(p.stat_aux.ptr + (int64_t)n * p.stat_aux.n_stride + (int64_t)oz * p.stat_aux.plane_elems() + pos +
                                 c * p.stat_aux.c_stride);
          }
          if (ok && c < p.out_c8) {
            if constexpr (kRes) {
              if (res_row && !use_ring) pre_res[c] = ldg16(res_row + c * p.res.c_stride);
            }
            if constexpr (kMaskAcc) {
              if (mask_row) pre_mask[c] = ldg16(mask_row + c * p.mask.c_stride);
              if (p.accumulate) pre_acc[c] = *reinterpret_cast<const uint4*>(out_row + c * p.out.c_stride);
            }
          }
        }
        if constexpr (kRing) {
          if (use_ring) {  // request the next block's vectors, then pick up this block's (requested one block ago)
            if (oz + kEgStep < zo1) {
              ring_issue(oz + kEgStep, rb ^ 1u);
              cp_async_wait<1>();
            } else {
              cp_async_wait<0>();
            }
            if (ok) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < p.out_c8) {
                  if constexpr (STAT == 2) pre_aux[c] = pre_ring[rb * 4 + c];
                  else pre_res[c] = pre_ring[rb * 4 + c];
                }
              }
            }
            rb ^= 1u;
          }
        }
        mbar_wait(&bar_acc_full[b], (full_mask >> b) & 1);
        full_mask ^= 1u << b;
        fence_after_sync();
#pragma unroll
        for (int c16 = 0; c16 < 5; ++c16) {  // NPo <= 80; unrolled so the prefetched vectors stay in registers
          if (c16 * 16 >= p.NPo) break;
          uint32_t v[16];
          tmem_ld16(trow + b * p.NPo + c16 * 16, v);
          tmem_ld_wait();
          if (c16 * 16 + 16 >= p.NPo) {  // last read of this block: zero it and hand it back to the MMA warp
            for (int c = 0; c < p.NPo; c += 16) tmem_st16_zero(trow + b * p.NPo + c);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(&bar_acc_empty[b]);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = c16 * 2 + h;
            if (ch >= p.out_c8 || !ok) continue;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[h * 8 + i]);
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] += __ldg(p.bias + ch * 8 + i);
            }
            if constexpr (kRes) {
              if (res_row) {
                float g[8];
                unpack8(ch < 4 ? pre_res[ch & 3] : ldg16(res_row + ch * p.res.c_stride), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += g[i];
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            bf16* dst = out_row + ch * p.out.c_stride;
            if constexpr (kMaskAcc) {
              if (mask_row) {
                float g[8];
                unpack8(ch < 4 ? pre_mask[ch & 3] : ldg16(mask_row + ch * p.mask.c_stride), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = g[i] > 0.f ? f[i] : 0.f;
              }
              if (p.accumulate) {
                float g[8];
                unpack8(ch < 4 ? pre_acc[ch & 3] : *reinterpret_cast<const uint4*>(dst), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += g[i];
              }
            }
            const uint4 pk = pack8(f);
            if constexpr (STAT != 0) {
              if (ch < 4) {  // statistics of the value as stored (bf16-rounded), like a separate pass over the tensor
                float r8[8], a8[8];
                unpack8(pk, r8);
                if constexpr (STAT == 2) unpack8(pre_aux[ch & 3], a8);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  st0[(ch & 3) * 8 + i] += r8[i];
                  st1[(ch & 3) * 8 + i] += r8[i] * (STAT == 2 ? a8[i] : r8[i]);
                }
              }
            }
            stg16(dst, pk);
          }
        }
      }
      if constexpr (STAT != 0) {
        // fold the 64 partial sums over the 32 lanes: after the five exchange rounds lane L holds entries 2L, 2L+1 of
        // [sum0[0..31], sum1[0..31]]
        float a[32], b[16], c[8], d[4], e[2];
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = (h16 ? st1[i] : st0[i]) + __shfl_xor_sync(0xffffffffu, h16 ? st0[i] : st1[i], 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = (h8 ? a[i + 16] : a[i]) + __shfl_xor_sync(0xffffffffu, h8 ? a[i] : a[i + 16], 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = (h4 ? b[i + 8] : b[i]) + __shfl_xor_sync(0xffffffffu, h4 ? b[i] : b[i + 8], 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = (h2 ? c[i + 4] : c[i]) + __shfl_xor_sync(0xffffffffu, h2 ? c[i] : c[i + 4], 2);
#pragma unroll
        for (int i = 0; i < 2; ++i) e[i] = (h1 ? d[i + 2] : d[i]) + __shfl_xor_sync(0xffffffffu, h1 ? d[i] : d[i + 2], 1);
        float2* slot = reinterpret_cast<float2*>(wslab + n * 64) + lane;  // always touched by this lane only
        float2 o = *slot;
        o.x += e[0];
        o.y += e[1];
        *slot = o;
#pragma unroll
        for (int i = 0; i < 32; ++i) st0[i] = st1[i] = 0.f;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == (LANES == 2 ? 2 : 1)) tmem_dealloc<512>(tmem_base_s);
  if constexpr (STAT != 0) {  // the 8 warp slabs of this CTA -> its CTA slab (fixed order)
    const int per = p.out.N * 64;
    const float* ws = p.stat_ws + (size_t)blockIdx.x * 8 * per;
    float* cs = p.stat_ws + (size_t)gridDim.x * 8 * per + (size_t)blockIdx.x * per;
    for (int i = tid; i < per; i += (int)blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += ws[(size_t)w * per + i];
      cs[i] = t;
    }
  }
}

// CTA slabs -> GroupNorm statistics (mode 1) or GroupNorm-backward reductions (mode 2); one block per sample
__global__ void __launch_bounds__(256) stat_finalize_kernel(const float* __restrict__ ws, int nslab, int N, int C, int G, double count,
                                                            float eps, const float* __restrict__ stats_in, float* __restrict__ out,
                                                            int mode) {
  __shared__ double part[4][64];
  __shared__ double tot[64];
  const int n = blockIdx.x, k = threadIdx.x & 63, q = threadIdx.x >> 6;
  const float* cs = ws + (size_t)nslab * 8 * N * 64 + (size_t)n * 64 + k;
  double t = 0;
  int s = q;
  for (; s + 28 < nslab; s += 32) {  // 8 independent loads in flight
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = cs[(size_t)(s + 4 * j) * N * 64];
#pragma unroll
    for (int j = 0; j < 8; ++j) t += (double)v[j];
  }
  for (; s < nslab; s += 4) t += (double)cs[(size_t)s * N * 64];
  part[q][k] = t;
  __syncthreads();
  if (threadIdx.x < 64) tot[k] = part[0][k] + part[1][k] + part[2][k] + part[3][k];
  __syncthreads();
  const int cpg = C / G;
  if (mode == 1) {
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      double s = 0, qq = 0;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += tot[c]; qq += tot[32 + c]; }
      const double m = count * cpg, mean = s / m;
      double var = qq / m - mean * mean;
      if (var < 0) var = 0;
      out[((size_t)n * G + g) * 2] = (float)mean;
      out[((size_t)n * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  } else if (threadIdx.x < C) {
    const int c = threadIdx.x, g = c / cpg;
    const double mean = stats_in[((size_t)n * G + g) * 2], rstd = stats_in[((size_t)n * G + g) * 2 + 1];
    out[((size_t)n * C + c) * 2] = (float)tot[c];                               // sum dy
    out[((size_t)n * C + c) * 2 + 1] = (float)(rstd * (tot[32 + c] - mean * tot[c]));  // sum dy * xhat
  }
}

struct Plan {
  uint32_t pre_off = 0;
  int KG, npass, PW, ntile, ZC, nzc, nstages, lanes;
  uint32_t stage_bytes, wbuf_bytes;
  size_t smem;
  bool ok;
};

constexpr size_t kPreRingBytes = 2 * 128 * 2 * 4 * 16;  // 2 lanes x 128 epilogue rows x 2 buffers x 4 chunks x 16 B

Plan make_plan(int K, int NPo, int Z, int X, int Y, bool pre_ring = false) {
  Plan pl{};
  pl.ok = false;
  if (K % 16 != 0 || NPo % 16 != 0 || NPo < 16 || 3 * NPo > 256) return pl;
  const int N3 = 3 * NPo;
  int KG = (9 * 32 * N3 * 2 <= 56 * 1024) ? 32 : 16;
  // K = 32 with a wide N (e.g. the paired space-to-depth dgrad, N3 = 192): a single pass needs only ONE weight buffer,
  // which then stays resident for the whole kernel instead of being re-fetched per (unit, pass)
  if (KG == 16 && K == 32 && (size_t)9 * 32 * N3 * 2 + 4 * (size_t)4 * (128 + 2 * (Y + 2) + 2) * 16 <= 220 * 1024) KG = 32;
  if (K < KG) KG = K;
  if (K % KG != 0) KG = 16;  // e.g. K = 48 (dgrad of the 45-channel regression conv): three passes of 16
  if (K % KG != 0) return pl;
  pl.KG = KG;
  pl.npass = K / KG;
  const int Yp = Y + 2;
  pl.PW = 128 + 2 * Yp + 2;
  pl.ntile = (X * Yp + 127) / 128;
  pl.ZC = 512 / NPo;
  if (pl.ZC > kMaxBlocks) pl.ZC = kMaxBlocks;
  if (pl.ZC > Z) pl.ZC = Z;
  if (pl.ZC < 2 && Z > 1) return pl;
  pl.nzc = (Z + pl.ZC - 1) / pl.ZC;
  pl.stage_bytes = (uint32_t)(KG / 8) * pl.PW * 16;
  pl.wbuf_bytes = (uint32_t)9 * KG * N3 * 2;
  const size_t wtotal = (size_t)(pl.npass > 1 ? 2 : 1) * pl.wbuf_bytes;
  size_t budget = 220 * 1024;
  // the residual / aux prefetch ring of the dual-lane kernel comes out of the stage budget (only where lanes == 2 is possible)
  const bool ring_ok = pre_ring && pl.npass == 1 && NPo <= 32 && Z >= 2 && ((Z + 1) / 2) * NPo <= 256;
  if (ring_ok) budget -= kPreRingBytes;
  if (wtotal + 2 * (size_t)pl.stage_bytes > budget) return pl;
  int S = (int)((budget - wtotal) / pl.stage_bytes);
  if (S > kMaxStages) S = kMaxStages;
  pl.nstages = S;
  pl.smem = wtotal + (size_t)S * pl.stage_bytes;
  pl.lanes = 1;
  // dual issue (two z-chunk lanes per CTA, see the kernel): single K pass, <= 32 result channels, both chunks' accumulators
  // in TMEM (2 * ceil(Z/2) * NPo <= 512 columns), at least two stages per lane
  static const bool no_dual = getenv("RTP_NO_DUAL") != nullptr;  // A/B switch
  const int zc2 = (Z + 1) / 2;
  if (!no_dual && pl.npass == 1 && NPo <= 32 && Z >= 2 && zc2 * NPo <= 256 && zc2 <= kMaxBlocks / 2 && S >= 4) {
    pl.lanes = 2;
    pl.ZC = zc2;
    pl.nzc = 2;
    pl.nstages = S & ~1;
    pl.smem = wtotal + (size_t)pl.nstages * pl.stage_bytes;
    if (ring_ok) {
      pl.pre_off = (uint32_t)pl.smem;
      pl.smem += kPreRingBytes;
    }
  }
  pl.ok = true;
  return pl;
}

int plan_units(const Plan& pl, int N) { return pl.lanes == 2 ? N * pl.ntile : N * pl.ntile * pl.nzc; }

}  // namespace

extern "C" int64_t rtp_conv_k3s1_smem_bytes(int32_t Cin, int32_t NPo, int32_t Z, int32_t X, int32_t Y) {
  Plan pl = make_plan(Cin, NPo, Z, X, Y);
  return pl.ok ? (int64_t)pl.smem : -1;
}

extern "C" int rtp_conv_k3s1(const rtp_conv_k3s1_desc* d, void* stream) {
  RTP_CHECK_ARG(d && d->in.ptr && d->out.ptr && d->w, "rtp_conv_k3s1: null argument");
  RTP_CHECK_ARG(d->in.N == d->out.N && d->in.Z == d->out.Z && d->in.X == d->out.X && d->in.Y == d->out.Y,
                "rtp_conv_k3s1: in/out geometry mismatch");
  RTP_CHECK_ARG(d->in.C8 * 8 >= d->Cin, "rtp_conv_k3s1: input has %d channels, K=%d", d->in.C8 * 8, d->Cin);
  RTP_CHECK_ARG(d->out_c8 >= 1 && d->out_c8 * 8 <= d->NPo + 7 && d->out_c8 <= d->out.C8, "rtp_conv_k3s1: bad out_c8");
  RTP_CHECK_ARG(d->in.c_stride == (int64_t)d->in.Z * (d->in.X + 2) * (d->in.Y + 2) * 8,
                "rtp_conv_k3s1: input planes must be contiguous per channel chunk");
  // The cp.async prefetch ring is OFF by default: measured in the step it is slower than the register prefetch (32->32:
  // 929 vs 1016 TFLOP/s, 21.25 vs 20.99 ms per step).  RTP_PRE_RING=1 enables it (A/B).  Restructuring the epilogue for it
  // did lower the register count of the statistics variants (157 -> 145 / 149), which is where the default path's gain over
  // the previous build (932-984 TFLOP/s) comes from.
  static const bool ring_on = getenv("RTP_PRE_RING") != nullptr;
  const bool want_ring = ring_on && ((d->stat_mode == 1 && d->res.ptr != nullptr) || d->stat_mode == 2);
  Plan pl = make_plan(d->Cin, d->NPo, d->in.Z, d->in.X, d->in.Y, want_ring);
  RTP_CHECK_ARG(pl.ok, "rtp_conv_k3s1: unsupported shape K=%d NPo=%d Z=%d X=%d Y=%d", d->Cin, d->NPo, d->in.Z, d->in.X, d->in.Y);
  K3 k;
  k.in = P8(d->in); k.out = P8(d->out); k.res = P8(d->res); k.mask = P8(d->mask);
  k.w = (const bf16*)d->w; k.bias = d->bias;
  k.NPo = d->NPo; k.N3 = 3 * d->NPo; k.out_c8 = d->out_c8; k.relu = d->relu; k.accumulate = d->accumulate;
  k.has_res = d->res.ptr != nullptr; k.has_mask = d->mask.ptr != nullptr;
  k.KG = pl.KG; k.npass = pl.npass; k.PW = pl.PW; k.ntile = pl.ntile; k.ZC = pl.ZC; k.nzc = pl.nzc;
  k.nunits = plan_units(pl, d->in.N);
  static int nsm = 0;
  if (!nsm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  k.split_from = 0x7fffffff;
  k.split_zh = 0;
  k.unit_list = d->unit_list;
  k.unit_count = d->unit_count;
  if (d->unit_list) RTP_CHECK_ARG(d->unit_count != nullptr, "rtp_conv_k3s1: unit_list without unit_count");
  static const bool no_split = getenv("RTP_NO_TAIL_SPLIT") != nullptr;  // A/B switch
  if (!no_split && !d->unit_list && ((pl.lanes == 1 && pl.nzc == 1 && d->in.Z >= 4 && d->in.Z % 2 == 0) || pl.lanes == 2)) {
    // lanes == 2: the two z-chunk lanes of a last-round unit go to two CTAs (a lone lane has the tensor pipe to itself)
    const int rem = k.nunits % nsm;
    if (k.nunits > nsm && rem > 0 && 2 * rem <= nsm) {  // e.g. 352 units on 148 SMs: 296 full + 112 half units
      k.split_from = k.nunits - rem;
      k.split_zh = d->in.Z / 2;
      k.nunits += rem;
    }
  }
  k.nstages = pl.nstages; k.stage_bytes = pl.stage_bytes; k.wbuf_bytes = pl.wbuf_bytes;
  k.pre_off = pl.lanes == 2 ? pl.pre_off : 0;
  k.wtap_bytes = (uint32_t)(pl.KG / 8) * k.N3 * 16;          // one tap's [KG/8][N3][8] slice
  k.wtap_stride = (uint32_t)(d->Cin / 8) * k.N3 * 16;        // distance between taps in the packed weights
  k.dbg = (long long*)d->debug;
  RTP_CHECK_ARG(d->stat_mode >= 0 && d->stat_mode <= 2, "rtp_conv_k3s1: bad stat_mode");
  if (d->stat_mode) {
    RTP_CHECK_ARG(d->stat_ws && d->out_c8 <= 4, "rtp_conv_k3s1: fused statistics need a workspace and <= 32 output channels");
    RTP_CHECK_ARG(!d->mask.ptr && !d->accumulate && (d->stat_mode == 1 || !d->res.ptr),
                  "rtp_conv_k3s1: stat_mode 1 supports bias/res/relu only, stat_mode 2 bias/relu only");
    if (d->stat_mode == 2)
      RTP_CHECK_ARG(d->stat_aux.ptr && d->stat_aux.N == d->out.N && d->stat_aux.Z == d->out.Z && d->stat_aux.X == d->out.X &&
                        d->stat_aux.Y == d->out.Y && d->stat_aux.C8 >= d->out_c8,
                    "rtp_conv_k3s1: stat_aux must have the output's geometry");
  }
  k.stat_aux = P8(d->stat_aux); k.stat_ws = d->stat_ws;
  RTP_CHECK_ARG(pl.npass <= 32, "rtp_conv_k3s1: too many K passes");
  for (int g = 0; g < 32; ++g) k.tapmask[g] = d->use_tap_mask ? (uint16_t)(d->tap_mask[g < pl.npass ? g * pl.KG / (d->Cin / d->tap_mask_groups) : 0] & 0x1FF) : 0x1FF;
  if (d->use_tap_mask) RTP_CHECK_ARG(d->tap_mask_groups >= 1 && d->tap_mask_groups <= 8 && d->Cin % d->tap_mask_groups == 0 &&
                                         (d->Cin / d->tap_mask_groups) % pl.KG == 0,
                                     "rtp_conv_k3s1: tap_mask_groups must split K into whole passes");
  const int ki = (pl.KG == 32 ? 1 : 0) + 2 * d->stat_mode + (pl.lanes == 2 ? 6 : 0);
  void (*kerns[12])(const K3) = {conv_k3s1_kernel<1, 0, 1>, conv_k3s1_kernel<2, 0, 1>, conv_k3s1_kernel<1, 1, 1>,
                                 conv_k3s1_kernel<2, 1, 1>, conv_k3s1_kernel<1, 2, 1>, conv_k3s1_kernel<2, 2, 1>,
                                 conv_k3s1_kernel<1, 0, 2>, conv_k3s1_kernel<2, 0, 2>, conv_k3s1_kernel<1, 1, 2>,
                                 conv_k3s1_kernel<2, 1, 2>, conv_k3s1_kernel<1, 2, 2>, conv_k3s1_kernel<2, 2, 2>};
  auto kern = kerns[ki];
  static size_t configured_dev[RTP_MAX_DEVICES][12];  /* the opt-in is per device */
  size_t* configured = configured_dev[rtp_current_device()];
  if (pl.smem > configured[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) { rtp_set_error("rtp_conv_k3s1: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured[ki] = pl.smem;
  }
  const int grid = k.nunits < nsm ? k.nunits : nsm;
  kern<<<grid, pl.lanes == 2 ? kThreads2 : kThreads, pl.smem, (cudaStream_t)stream>>>(k);
  RTP_LAUNCH_CHECK();
}

extern "C" int64_t rtp_conv_k3s1_stat_ws_bytes(int32_t N) {
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  return (int64_t)nsm * 9 * N * 64 * 4;
}

extern "C" int32_t rtp_conv_k3s1_num_ctas(int32_t Cin, int32_t NPo, int32_t N, int32_t Z, int32_t X, int32_t Y) {
  Plan pl = make_plan(Cin, NPo, Z, X, Y);
  if (!pl.ok) return -1;
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int nunits = plan_units(pl, N);
  return nunits < nsm ? nunits : nsm;
}

extern "C" int rtp_conv_k3s1_stat_finalize(const float* stat_ws, int32_t nctas, int32_t mode, int32_t N, int32_t C, int32_t G,
                                           int64_t voxels, float eps, const float* stats_in, float* out, void* stream) {
  RTP_CHECK_ARG(stat_ws && out && nctas >= 1 && N >= 1 && C >= 1 && C <= 32 && G >= 1 && C % G == 0 && (mode == 1 || mode == 2),
                "rtp_conv_k3s1_stat_finalize: bad arguments");
  RTP_CHECK_ARG(mode == 1 || stats_in, "rtp_conv_k3s1_stat_finalize: mode 2 needs the forward statistics");
  stat_finalize_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(stat_ws, nctas, N, C, G, (double)voxels, eps, stats_in, out, mode);
  RTP_LAUNCH_CHECK();
}

int rtp_k3s1_set_carveout(int pct) {  // see rtp_set_shared_carveout (layout.cu)
  return (int)cudaFuncSetAttribute((const void*)stat_finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

namespace {
// one block of 1024 threads: flags in shared memory, then an ordered compaction (a few thousand units at most)
__global__ void __launch_bounds__(1024) active_units_kernel(const int64_t* __restrict__ ind, int N, int M, int X, int Y, int radius,
                                                            int ntile, int* __restrict__ list, int* __restrict__ count) {
  extern __shared__ int au_flags[];
  const int total = N * ntile, Yp = Y + 2;
  for (int i = threadIdx.x; i < total; i += blockDim.x) au_flags[i] = 0;
  __syncthreads();
  const int side = 2 * radius + 1;
  const int items = N * M * side * side;
  const int64_t YX = (int64_t)Y * X;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int dy = i % side - radius, dx = (i / side) % side - radius;
    const int j = i / (side * side);  // n * M + target
    const int64_t id = ind[j];
    const int r = (int)(id % YX);
    const int y = r / X + dy, x = r % X + dx;
    if (x < 0 || x >= X || y < 0 || y >= Y) continue;
    const int q = (x + 1) * Yp + (y + 1);
    au_flags[(j / M) * ntile + (q - Yp) / 128] = 1;  // benign race: every writer stores 1
  }
  __syncthreads();
  // ordered compaction: thread t owns the flags [t * per, (t + 1) * per); block-wide exclusive scan of the per-thread counts
  // (one thread walking all flags took 17 us at the bench shape, four times per step on the critical path)
  __shared__ int au_warp[32];
  const int per = (total + (int)blockDim.x - 1) / (int)blockDim.x;
  const int b0 = min(total, (int)threadIdx.x * per), b1 = min(total, b0 + per);
  int c = 0;
  for (int i = b0; i < b1; ++i) c += au_flags[i] != 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) au_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < (int)(blockDim.x >> 5) ? au_warp[lane] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) sc += u;
    }
    au_warp[lane] = sc - v;  // exclusive prefix of the warp totals
    if (lane == 31) *count = sc;
  }
  __syncthreads();
  int pos = au_warp[wid] + incl - c;
  for (int i = b0; i < b1; ++i)
    if (au_flags[i]) list[pos++] = i;
}
}  // namespace

extern "C" int rtp_active_units(const int64_t* ind, int32_t N, int32_t M, int32_t Z, int32_t X, int32_t Y, int32_t radius,
                                int32_t* unit_list, int32_t* unit_count, void* stream) {
  RTP_CHECK_ARG(ind && unit_list && unit_count && N >= 1 && M >= 1 && Z >= 1 && X >= 1 && Y >= 1 && radius >= 0 && radius <= 8,
                "rtp_active_units: bad arguments");
  const int ntile = (X * (Y + 2) + 127) / 128;
  RTP_CHECK_ARG((int64_t)N * ntile <= 11000, "rtp_active_units: too many units for one block's shared memory");
  active_units_kernel<<<1, 1024, (size_t)N * ntile * sizeof(int), (cudaStream_t)stream>>>(ind, N, M, X, Y, radius, ntile, unit_list,
                                                                                          unit_count);
  RTP_LAUNCH_CHECK();
}
