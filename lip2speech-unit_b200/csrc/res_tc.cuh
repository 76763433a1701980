// A whole ResBlock1 (speech-resynthesis/models.py:16-45) in one kernel for the narrow stages (C <= 64):
//
//     for d in dilations:  x = x + c2_d( lrelu( c1_d( lrelu(x) ) ) )
//
// One step per launch (pair_tc.cuh) is HBM bound there: every step re-reads the activated input and the
// fp32 residual and re-writes both.  Here the residual stream never leaves the SM between the steps:
//
//   X  (TMEM, fp32, msub x C columns)   the residual stream of the tile.  It is written once from global
//                                       memory (tcgen05.st, as x + b2 of the first step) and after that only by
//                                       the c2 MMAs, which ACCUMULATE onto it; the bias of the next c2 is added
//                                       by the epilogue warps while the next c1 runs.
//   D1 (TMEM, fp32, msub x C columns)   accumulator of c1, pre-loaded with the bias of the next c1 while c2 runs,
//                                       so that phases A and B are TMEM load -> leaky-ReLU -> bf16 -> shared store
//   S  (shared, bf16, (MT + 2 pad) rows in the K-major swizzled operand layout)
//                                       lrelu(x), then lrelu(c1 + b1), then lrelu(x') ... each overwrites the
//                                       previous one once the MMAs that read it have retired.
//
// A tile is MT = 128 * msub rows; every conv reads row-shifted views of S, so rows closer than the summed halo
// H = sum_d (d + 1)(k - 1)/2 to a tile edge are wrong at the end and are simply not stored: items advance by
// R = MT - 2H rows.  Rows outside [0, L) are forced to zero in S before every conv (the per-layer zero padding
// of the reference).
//
//   warp 0      TMA producer: the weight stages of the 2 n_dil convs, in order, through a ring; L2 prefetch of the
//               next item's x / branch-sum rows
//   warp 1      MMA issuer:   wait s_full -> c1 -> commit d_full -> wait s_full -> c2 (onto X) -> commit d_full
//   warps 2..9  load x (global, one row per lane -> X, S), phase A (D1 -> S), phase B (X -> S), and after the last
//               step the transposed global epilogue of conv_tc.cuh (branch sum / mean / activated copy)
//
// MMA and epilogue phases of one CTA alternate; two co-resident CTAs per SM overlap them.
// Rounding: x + sum of products is formed in the tensor core's fp32 accumulator and the biases enter at different
// points than in the step-by-step kernels, so results differ from those in the last fp32 bits (tests compare the
// two with a tolerance; tile size and CTAs per SM do not change a bit).
#pragma once
#include "pair_tc.cuh"

namespace l2s {

constexpr int kResMaxDil = 4;

struct ResGeom {
  int c, k, n_dil;
  int dil[kResMaxDil], h1[kResMaxDil];
  int h2, h_tot, pad;
  int rb, k16, lc;          // lc = log2(c)
  int msub, mt, r_out, m_items, total_items;
  int s_bytes;
  int tb, n_tstages, bstage_bytes, sb;
  int tmem_cols, cw, dual, tile_words;
  int ne;                   // epilogue warps: 8, or 4 in the four-CTAs-per-SM plans (one warp per TMEM lane quadrant)
  int iss2;                 // 1: a second MMA-issuing warp (index 2 + ne) takes the upper half of the accumulators
  int ctas_per_sm;          // 1, 2 (dual) or 4 (quad)
  int cg2;                  // 1: CTA pairs (cluster of 2) issue cta_group::2 MMAs; a CTA keeps half of every weight stage
  uint32_t idesc;
  int smem_bytes;
  // skewed schedule (resq_tc.cuh): two S slabs, per-granule barriers, weight-stage groups walked granule by granule
  int skew;                 // 1: run by resq_tc_kernel
  int gran, ng;             // granule = gran 128-row accumulators (32 TMEM columns at least), ng granules per tile
  int gh, gt, head_fwd;     // first gh / last gt weight stages of a conv are walked granule-outer; head_fwd: the head group
                            // holds a tap right of the centre (its MMAs on granule i read S rows of granule i + 1)
};

struct ResParams {
  ConvParams c;                        // output epilogue: bias = zeros (X already holds every bias), acc_in, out_raw, out_act, div, slope
  const float* x;                      // fp32 residual stream entering the block, [B][L][C]
  const float* bias1[kResMaxDil];      // c1 biases
  const float* bias2[kResMaxDil];      // c2 biases
  ResGeom g;
  unsigned long long* span;            // debug: [0] min CTA start, [1] max CTA end (globaltimer)
  long long* trace;                    // debug: globaltimer stamps of CTA 0, [0..127] epilogue warp 2, [128..255] MMA warp
};
#define L2S_RTRACE(base, n)                                                          \
  do {                                                                               \
    if (P.trace && blockIdx.x == 0 && lane == 0 && (n) < 128) P.trace[(base) + (n)++] = gtime(); \
  } while (0)

struct ResMaps {
  CUtensorMap w[2 * kResMaxDil];       // c1_0, c2_0, c1_1, c2_1, ...
};

// The epilogue warps walk the tile's TMEM region (X or D1: msub accumulators x C columns) in 32-column units; the
// two warps of a lane quadrant alternate.  For C >= 32 a unit is 32 channels of one accumulator (one tile row per
// lane); for C = 16 it is two neighbouring accumulators (two tile rows per lane, 128 apart).
struct ResLane {
  int quad, half, lane;
  int ustep;     // units advance by this much per warp: 2 when two warps share a lane quadrant, 1 when one warp owns it
  int sw;        // this lane's swizzle term of the S slots (rows advance by multiples of 8 between units)
};

__device__ __forceinline__ void res_unit_pos(const ResGeom& g, const ResLane& w, int u, int& row, int& c0) {
  const int flat = 32 * u;
  const int s = flat >> g.lc;
  c0 = flat & (g.c - 1);
  row = s * 128 + w.quad * 32 + w.lane;
}

// 32 fp32 values of one unit -> leaky-ReLU(0.1) -> bf16 -> S (zero where the row lies outside [0, L) when EDGE).
template <bool EDGE>
__device__ __forceinline__ void res_store_act(const ResGeom& g, const ResLane& w, uint8_t* slab, const uint32_t (&r)[32], int row,
                                              int c0, int t_row0, int lin) {
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(0.1f);   // LRELU_SLOPE, models.py:13,38
  const bool c16 = g.c == 16;
#pragma unroll
  for (int e = 0; e < 4; ++e) {                               // 8 columns = one 16-byte slot
    const int rr = (c16 && e >= 2) ? row + 128 : row;         // C = 16: columns 16..31 belong to the next accumulator
    const int ch = c16 ? (e & 1) * 8 : c0 + e * 8;
    uint4 pk;
    pk.x = lrelu_bf16x2(__uint_as_float(r[8 * e + 0]), __uint_as_float(r[8 * e + 1]), slope2);
    pk.y = lrelu_bf16x2(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]), slope2);
    pk.z = lrelu_bf16x2(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]), slope2);
    pk.w = lrelu_bf16x2(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]), slope2);
    if (EDGE) {
      const int t = t_row0 + rr;
      if (t < 0 || t >= lin) pk = make_uint4(0u, 0u, 0u, 0u); // the conv's zero padding
    }
    *reinterpret_cast<uint4*>(slab + (size_t)(g.pad + rr) * g.rb + ((((ch >> 3) ^ w.sw)) << 4)) = pk;
  }
}

// Phase A / B: TMEM (D1 or X, already holding the bias) -> S.  Nothing but the TMEM load, the conversion and the
// shared-memory stores sits between the MMAs that produced the values and the MMAs that consume them.
template <bool EDGE>
__device__ __forceinline__ void res_phase(const ResGeom& g, const ResLane& w, uint8_t* slab, uint32_t t_quad, int t_row0, int lin,
                                          long long* tr = nullptr) {
  const int n_units = (g.msub * g.c) >> 5;
  for (int u = w.half; u < n_units; u += w.ustep) {
    uint32_t r[32];
    if (tr && w.lane == 0) tr[0] = gtime();
    tmem_ld32(t_quad + (uint32_t)(32 * u), r);
    tmem_ld_wait();
    if (tr && w.lane == 0) tr[1] = gtime();
    int row, c0;
    res_unit_pos(g, w, u, row, c0);
    res_store_act<EDGE>(g, w, slab, r, row, c0, t_row0, lin);
    if (tr && w.lane == 0) { tr[2] = gtime(); tr += 3; }
  }
}

// Off the critical path (the tensor core is busy with the conv that was just released):
// D1 <- bias of the next c1, so that its MMAs accumulate onto the bias and phase A needs no additions.
__device__ __forceinline__ void res_prebias_d1(const ResGeom& g, const ResLane& w, uint32_t d1_quad, const float* sbias) {
  const int n_units = (g.msub * g.c) >> 5;
  const int cmask = g.c - 1;
  for (int u = w.half; u < n_units; u += w.ustep) {
    const int c0 = (32 * u) & cmask;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bq = *reinterpret_cast<const float4*>(sbias + ((c0 + 4 * j) & cmask));
      r[4 * j] = __float_as_uint(bq.x); r[4 * j + 1] = __float_as_uint(bq.y);
      r[4 * j + 2] = __float_as_uint(bq.z); r[4 * j + 3] = __float_as_uint(bq.w);
    }
    tmem_st32(d1_quad + (uint32_t)(32 * u), r);
  }
  tmem_st_wait();
}
// X <- X + bias of the next c2 (X is idle between phase B and the next c2).
__device__ __forceinline__ void res_addbias_x(const ResGeom& g, const ResLane& w, uint32_t x_quad, const float* sbias) {
  const int n_units = (g.msub * g.c) >> 5;
  const int cmask = g.c - 1;
  for (int u = w.half; u < n_units; u += w.ustep) {
    const int c0 = (32 * u) & cmask;
    uint32_t r[32];
    tmem_ld32(x_quad + (uint32_t)(32 * u), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bq = *reinterpret_cast<const float4*>(sbias + ((c0 + 4 * j) & cmask));
      r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + bq.x);
      r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + bq.y);
      r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + bq.z);
      r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + bq.w);
    }
    tmem_st32(x_quad + (uint32_t)(32 * u), r);
  }
  tmem_st_wait();
}

// x (global fp32) -> one row per lane -> S = lrelu(x), X = x + b2_0.  Every lane reads its own row(s) with 16-byte
// loads: a row is 64..256 contiguous bytes, so the sectors it touches are fully used by this lane's consecutive
// loads (L1 serves the second half), and no shared-memory transpose is needed on the way in.
__device__ __forceinline__ void res_load_x(const ResParams& P, const ResLane& w, uint8_t* slab, uint32_t x_quad, int b,
                                           int t_row0, const float* sbias2, int lin) {
  const ResGeom& g = P.g;   // lin: rows of the utterance (0 for the dummy item of an odd CTA pair: everything reads as zero)
  const int n_units = (g.msub * g.c) >> 5;
  const int cmask = g.c - 1;
  const bool c16 = g.c == 16;
  const float* xb = P.x + (long long)b * lin * g.c;
  // all of this lane's rows into L1 first (one prefetch per 128-byte line, no registers held), so that the unit
  // loop below pays the L2 latency once instead of once per unit
  for (int u = w.half + w.ustep; u < n_units; u += w.ustep) {
    int row, c0;
    res_unit_pos(g, w, u, row, c0);
    const int t = t_row0 + row;
    if (t >= 0 && t < lin) prefetch_l1(xb + (long long)t * g.c + c0);
    if (c16 && t + 128 >= 0 && t + 128 < lin) prefetch_l1(xb + (long long)(t + 128) * g.c);
  }
  for (int u = w.half; u < n_units; u += w.ustep) {
    int row, c0;
    res_unit_pos(g, w, u, row, c0);
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = (c16 && j >= 4) ? row + 128 : row;
      const int ch = c16 ? (j & 3) * 4 : c0 + 4 * j;
      const int t = t_row0 + rr;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t >= 0 && t < lin) q = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * g.c + ch));
      r[4 * j] = __float_as_uint(q.x); r[4 * j + 1] = __float_as_uint(q.y);
      r[4 * j + 2] = __float_as_uint(q.z); r[4 * j + 3] = __float_as_uint(q.w);
    }
    res_store_act<false>(g, w, slab, r, row, c0, t_row0, lin);   // rows outside [0, L) were loaded as zeros
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bq = *reinterpret_cast<const float4*>(sbias2 + ((c0 + 4 * j) & cmask));
      r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + bq.x);
      r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + bq.y);
      r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + bq.z);
      r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + bq.w);
    }
    tmem_st32(x_quad + (uint32_t)(32 * u), r);
  }
  tmem_st_wait();
}

// Output of the block: X (complete, biases included) -> transposed epilogue of conv_tc.cuh (branch sum / mean /
// activated copy).  Chunks that hold no output row (the halo at both tile ends) are skipped, and the branch-sum
// loads of the next chunk are in flight while the current one is transposed and stored.
template <int CW, int MODE, bool PIPE>
__device__ __forceinline__ void res_output(const ConvParams& p, float* tile, uint32_t t_base, int b, int t_row0, int row_lo,
                                           int row_lim, int msub, int c, const ResLane& w, uint64_t* bar, uint32_t parity) {
  constexpr int LPR = CW / 4;
  constexpr uint32_t kAll = (1u << (CW / 4)) - 1u;
  const int crow = w.lane / LPR, c4 = w.lane % LPR;
  float4* tile4 = reinterpret_cast<float4*>(tile);
  const int cps = c / CW;
  const int cps_sh = cps == 4 ? 2 : (cps == 2 ? 1 : 0);
  const int n_chunks = msub << cps_sh;
  // Ownership follows the 32-column units of res_load_x / res_phase (a warp that finishes early starts writing the
  // next item's X while its neighbour may still be reading this one): unit u = half, half + ustep, ... is chunk u for
  // 32-column chunks and chunks 2u, 2u + 1 for 16-column chunks.
  const int first = CW == 32 ? w.half : 2 * w.half;
  auto next_after = [&](int idx) { return CW == 32 ? idx + w.ustep : ((idx & 1) ? idx + 2 * w.ustep - 1 : idx + 1); };
  auto skip = [&](int idx) {   // first owned chunk at or after idx that holds an output row
    while (idx < n_chunks) {
      const int r0 = t_row0 + (idx >> cps_sh) * 128 + w.quad * 32;
      if (r0 < row_lim && r0 + 32 > row_lo) break;
      idx = next_after(idx);
    }
    return idx;
  };
  auto locate = [&](int idx) {
    const int s_ = idx >> cps_sh, cc_ = idx & (cps - 1);
    return epi_locate<CW>(p, t_base + (uint32_t)(s_ * c + cc_ * CW), b, t_row0 + s_ * 128 + w.quad * 32, cc_ * CW, crow, c4, row_lim,
                          row_lo);
  };
  // MODE & kEpiRes: a SECOND branch output to add (p.res: the ResBlock of another kernel-size branch that ran in parallel with
  // the one whose running sum is p.acc_in).  epi_finish adds it before the running sum: ((x_0 + x_1) + x_2) / 3, the
  // association of the reference's `xs += resblock(x)` loop and of the serial branch chain.
  auto finish = [&](const EpiChunk& ch, const float4 (&av)[CW / 4]) {
    float4 rv[CW / 4] = {};
    epi_load_res<CW, MODE>(p, ch, rv);       // in flight while the accumulator is staged
    epi_stage<CW>(tile4, ch.taddr, w.lane);
    __syncwarp();
    if (__all_sync(0xffffffffu, ch.okmask == kAll)) epi_finish<CW, MODE, true>(p, ch, tile4, rv, av, crow, c4);
    else epi_finish<CW, MODE, false>(p, ch, tile4, rv, av, crow, c4);
    __syncwarp();   // the tile is rewritten by the next chunk
  };
  int idx = skip(first);
  if constexpr (PIPE) {
    // 168-register budget: the next chunk's branch-sum values wait in registers
    EpiChunk ca{}, cb{};
    float4 ava[CW / 4], avb[CW / 4];
    if (idx < n_chunks) { ca = locate(idx); epi_load_acc<CW, MODE>(p, ca, ava); }
    if (bar) mbar_wait(bar, parity);            // nullptr: the caller has already waited for the accumulators
    tc_fence_after();
    while (idx < n_chunks) {
      const int idx2 = skip(next_after(idx));
      if (idx2 < n_chunks) { cb = locate(idx2); epi_load_acc<CW, MODE>(p, cb, avb); }
      finish(ca, ava);
      ca = cb;
#pragma unroll
      for (int i = 0; i < CW / 4; ++i) ava[i] = avb[i];
      idx = idx2;
    }
  } else {
    // 80-register budget: the next chunk's branch-sum lines are pulled into L1 instead (no registers held)
    constexpr int RPI = 32 / LPR, ITERS = 32 / RPI;
    auto prefetch_acc = [&](int i2) {
      if constexpr ((MODE & kEpiAcc) != 0) {
        if (p.acc_in && i2 < n_chunks && c4 == 0) {       // one lane per row: a row of the chunk is at most one line
          const int s_ = i2 >> cps_sh, cc_ = i2 & (cps - 1);
          const int qa = t_row0 + s_ * 128 + w.quad * 32 + crow;
#pragma unroll
          for (int i = 0; i < ITERS; ++i) {
            const int q = qa + i * RPI;
            if (q >= row_lo && q < row_lim) {
              prefetch_l1(p.acc_in + ((long long)b * p.lin + q) * p.ntot + cc_ * CW);
              if constexpr ((MODE & kEpiRes) != 0) prefetch_l1(p.res + ((long long)b * p.lin + q) * p.ntot + cc_ * CW);
            }
          }
        }
      }
    };
    prefetch_acc(idx);
    if (bar) mbar_wait(bar, parity);
    tc_fence_after();
    while (idx < n_chunks) {
      const EpiChunk ca = locate(idx);
      float4 ava[CW / 4];
      epi_load_acc<CW, MODE>(p, ca, ava);
      idx = skip(next_after(idx));
      prefetch_acc(idx);
      finish(ca, ava);
    }
  }
}

// MMAs of one weight stage (taps tap0 .. tap0 + t_end) for all msub accumulators.  Templated on C so that the
// instruction descriptor and every descriptor stride are immediates: with run-time values the compiler re-loaded
// them from the constant bank next to every UTCHMMA, and those dependent loads (not the tensor core) set the pace
// of these small N = C MMAs.
template <int C, bool CG2 = false>
__device__ __forceinline__ void res_issue_stage(bool leader, int msub, uint32_t desc_hi, uint32_t a_tap0_lo, uint32_t tap_step,
                                                uint32_t b_lo, int tap0, int t_end, uint32_t d_base) {
  constexpr int K16 = C / 16;
  constexpr uint32_t kRb = 2 * C;
  constexpr uint32_t kSubStep = (128u * kRb) >> 4;
  constexpr uint32_t kTapW = ((uint32_t)(CG2 ? C / 2 : C) * kRb) >> 4;      // CTA pair: a CTA holds half of the weight rows
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)C >> 3) << 17) | (((CG2 ? 256u : 128u) >> 4) << 24);
  uint32_t a_tap = a_tap0_lo + (uint32_t)tap0 * tap_step;
  for (int t = 0; t < t_end; ++t, b_lo += kTapW, a_tap += tap_step) {
    uint32_t a_sub = a_tap;
    uint32_t d_addr = d_base;
    for (int s = 0; s < msub; ++s, a_sub += kSubStep, d_addr += (uint32_t)C) {
#pragma unroll
      for (int k = 0; k < K16; ++k) {
        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_sub + 2u * k);
        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
        if (leader) {                                         // both convs accumulate: D1 starts as the bias, X as x
          if constexpr (CG2) umma_bf16_cg2(d_addr, da, db, kIdesc, 1u); else umma_bf16(d_addr, da, db, kIdesc, 1u);
        }
      }
    }
  }
}

// CG2: the grid is launched as CTA pairs (cluster of 2) walking items 2j, 2j + 1 in lockstep; the leader CTA's MMA
// thread issues cta_group::2 MMAs (M = 256: both tiles), each CTA loads HALF of every weight stage and keeps its own
// X / D1 / S; barriers the MMA thread waits on live in the leader and collect both CTAs' arrivals, its commits are
// multicast to both.  (Own instantiation: a kernel that contains cta_group::2 code cannot be launched without a cluster.)
// WIDE: one CTA per SM with SIXTEEN epilogue warps (four per TMEM lane quadrant, 576 threads, 96 registers: 18 warps put five on one scheduler, whose 16 K registers then bound the count): MMA and
// epilogue phases of a one-CTA-per-SM plan strictly alternate, so halving every phase's latency shortens the tile's chain.
template <int MODE, bool DUAL, bool CG2 = false, bool WIDE = false>
__global__ void __maxnreg__(DUAL ? 80 : (WIDE ? 96 : 168))
res_tc_kernel(const __grid_constant__ ResMaps maps, const ResParams P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = P.c;
  const ResGeom& g = P.g;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slab = smem;
  uint8_t* stageB = smem + (size_t)g.s_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageB + (size_t)g.sb * g.bstage_bytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + kTcMaxStagesB;
  uint64_t* s_full = b_empty + kTcMaxStagesB;   // S (and X) ready for the next conv: one arrival per epilogue warp
  uint64_t* d_full = s_full + 1;                // the conv's MMAs have retired (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
  float* sbias = reinterpret_cast<float*>(bars + 24);                       // [2 n_dil][C]: b1_s, b2_s
  float* epi_tiles = sbias + 2 * kResMaxDil * 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (P.span && threadIdx.x == 0) atomicMin(&P.span[0], (unsigned long long)gtime());

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2 * g.n_dil; ++i) tma_prefetch_desc(&maps.w[i]);
    const uint32_t n_iss = g.iss2 ? 2u : 1u;               // every issuing warp commits to b_empty / d_full
    for (int i = 0; i < g.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], n_iss); }
    mbar_init(s_full, (uint32_t)(CG2 ? 2 * g.ne : g.ne));
    mbar_init(d_full, n_iss);
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(tmem_slot, (uint32_t)g.tmem_cols); else tmem_alloc_dyn(tmem_slot, (uint32_t)g.tmem_cols); }
  if (warp >= 2) {
    // the pad rows above and below the tile are read by the outer taps but never written: zero them once
    const int pad_bytes = g.pad * g.rb;
    uint8_t* lo = slab;
    uint8_t* hi = slab + (size_t)(g.pad + g.mt) * g.rb;
    for (int o = (threadIdx.x - 64) * 16; o < pad_bytes; o += ((int)blockDim.x - 64) * 16) {
      *reinterpret_cast<uint4*>(lo + o) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(hi + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = threadIdx.x - 64; i < 2 * g.n_dil * g.c; i += (int)blockDim.x - 64) {
      const int cv = i / g.c, ch = i - cv * g.c;
      sbias[cv * 64 + ch] = (cv & 1) ? P.bias2[cv >> 1][ch] : P.bias1[cv >> 1][ch];
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();        // the partner's barriers exist before anything is signalled to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = g.msub * g.c;            // X at [0, acc_cols), D1 at [acc_cols, 2 acc_cols)
  // item walk: a CTA pair advances in lockstep (pair j handles items 2j + rank); an odd total leaves a dummy item
  const int crank = CG2 ? (int)cluster_ctarank() : 0;
  const int walkers = CG2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int walk0 = CG2 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int walk_n = CG2 ? (g.total_items + 1) / 2 : g.total_items;
  auto item_of = [&](int wk) { return CG2 ? 2 * wk + crank : wk; };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (weights only)
    const bool leader = elect_one();
    int ib = 0;
    uint32_t pb = 0;
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      // The next item's x (and branch-sum) rows are one contiguous range: pull them into L2 now, a whole item
      // ahead, so that the load phase and the output epilogue see L2 latency instead of DRAM latency.
      const int nxt = wk + walkers < walk_n ? item_of(wk + walkers) : g.total_items;
      if (leader && nxt < g.total_items) {
        const int nb = nxt / g.m_items;
        const int nq = (nxt - nb * g.m_items) * g.r_out;
        const int lo = max(nq - g.h_tot, 0), hi = min(nq - g.h_tot + g.mt, p.lin);
        const long long e0 = ((long long)nb * p.lin + lo) * g.c;
        const uint32_t bytes = (uint32_t)((hi - lo) * g.c * 4);
        for (uint32_t o = 0; o < bytes; o += 16384u)
          bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(P.x + e0) + o, min(16384u, bytes - o));
        if ((MODE & kEpiAcc) != 0 && p.acc_in) {
          const int olo = max(nq, 0), ohi = min(nq + g.r_out, p.lin);
          const long long a0 = ((long long)nb * p.lin + olo) * g.c;
          const uint32_t ab = (uint32_t)((ohi - olo) * g.c * 4);
          for (uint32_t o = 0; o < ab; o += 16384u)
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.acc_in + a0) + o, min(16384u, ab - o));
        }
      }
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if (leader) {
            if constexpr (CG2) {   // this CTA's half of the stage; the bytes of both halves are counted on the leader's barrier
              if (crank == 0) mbar_expect_tx(&b_full[ib], 2u * (uint32_t)g.bstage_bytes);
              tma_load_3d_cg2(stageB + (size_t)ib * g.bstage_bytes, &maps.w[cv], &b_full[ib], 0, crank * (g.c / 2), ts * g.tb);
            } else {
              mbar_expect_tx(&b_full[ib], (uint32_t)g.bstage_bytes);
              tma_load_3d(stageB + (size_t)ib * g.bstage_bytes, &maps.w[cv], &b_full[ib], 0, 0, ts * g.tb);
            }
          }
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || (g.iss2 && warp == 2 + g.ne)) {
    // -------------------------------------------------------------- MMA issuer(s)
    // iss2: warp 1 issues for accumulators [0, ceil(msub / 2)), warp 2 + ne for the rest; both wait for the same s_full /
    // b_full completions and both commit to b_empty / d_full (count 2).  Every accumulator still sees its MMAs in the same order.
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template((uint32_t)g.rb);
    const uint32_t desc_hi = (uint32_t)(tmpl >> 32);
    const uint32_t desc_lo_fixed = (uint32_t)tmpl;
    const uint32_t row_step = (uint32_t)g.rb >> 4;
    const int acc_first = (g.iss2 && warp != 1) ? (g.msub + 1) / 2 : 0;
    const int acc_count = g.iss2 ? (warp == 1 ? (g.msub + 1) / 2 : g.msub / 2) : g.msub;
    const uint32_t s_lo = (desc_lo_fixed | ((smem_u32(slab) & 0x3FFFFu) >> 4)) + (uint32_t)acc_first * ((128u * (uint32_t)g.rb) >> 4);
    int ib = 0;
    uint32_t pb = 0, ps = 0;
    int ntr = warp == 1 ? 0 : 128;     // trace stamps: the first issuer only
    const bool count_waits = P.trace && blockIdx.x == 0 && warp == 1;   // trace build: cycles spent waiting for weights / for the epilogue
    long long w_b = 0, w_s = 0;
    const long long c_start = count_waits ? clock64() : 0;
    auto commit = [&](uint64_t* bar) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); };
    for (int wk = (CG2 && crank != 0) ? walk_n : walk0; wk < walk_n; wk += walkers) {   // CTA pair: the leader issues for both
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        const int st = cv >> 1;
        const bool second = (cv & 1) != 0;
        const int dil = second ? 1 : g.dil[st];
        const int halo = second ? g.h2 : g.h1[st];
        L2S_RTRACE(128, ntr);
        {
          const long long c0 = count_waits ? clock64() : 0;
          mbar_wait(s_full, ps);
          if (count_waits) w_s += clock64() - c0;
        }
        ps ^= 1u;
        tc_fence_after();
        L2S_RTRACE(128, ntr);
        const uint32_t a_tap0 = s_lo + (uint32_t)(g.pad - halo) * row_step;
        const uint32_t d_base = (second ? tmem_base : tmem_base + (uint32_t)acc_cols) + (uint32_t)(acc_first * g.c);   // c2 accumulates onto X, c1 onto its bias
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          {
            const long long c0 = count_waits ? clock64() : 0;
            mbar_wait(&b_full[ib], pb);
            if (count_waits) w_b += clock64() - c0;
          }
          tc_fence_after();
          const uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * g.bstage_bytes) & 0x3FFFFu) >> 4);
          const int t_end = min(g.tb, g.k - ts * g.tb);
          if (g.c == 64) res_issue_stage<64, CG2>(leader, acc_count, desc_hi, a_tap0, (uint32_t)dil * row_step, b_lo, ts * g.tb, t_end, d_base);
          else if (g.c == 32) res_issue_stage<32, CG2>(leader, acc_count, desc_hi, a_tap0, (uint32_t)dil * row_step, b_lo, ts * g.tb, t_end, d_base);
          else res_issue_stage<16, CG2>(leader, acc_count, desc_hi, a_tap0, (uint32_t)dil * row_step, b_lo, ts * g.tb, t_end, d_base);
          if (leader) commit(&b_empty[ib]);
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
        if (leader) commit(d_full);
        L2S_RTRACE(128, ntr);
      }
    }
    if (count_waits && lane == 0) { P.trace[500] = w_b; P.trace[501] = w_s; P.trace[502] = clock64() - c_start; }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    constexpr int CW = DUAL ? 16 : 32;
    ResLane w;
    w.quad = warp & 3;
    w.half = (warp - 2) >> 2;
    w.ustep = g.ne >> 2;
    w.lane = lane;
    w.sw = g.rb == 128 ? (lane & 7) : (g.rb == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
    float* tile = epi_tiles + (size_t)(warp - 2) * g.tile_words;
    const uint32_t x_quad = tmem_base + ((uint32_t)(w.quad * 32) << 16);
    const uint32_t d1_quad = x_quad + (uint32_t)acc_cols;
    uint32_t pd = 0;
    int ntr = warp == 2 ? 0 : 128;
    auto publish = [&]() {               // S / X / D1 written by this warp are visible to the tensor core
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(s_full, 0u, (uint32_t)crank); else mbar_arrive(s_full); }   // pair: the leader's MMA thread waits
    };
    res_prebias_d1(g, w, d1_quad, sbias);            // bias of the first c1
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      const int item = item_of(wk);
      const bool dummy = item >= g.total_items;    // odd item count: the pair's last partner computes on zeros and stores nothing
      const int b = dummy ? 0 : item / g.m_items;
      const int mi = dummy ? 0 : item - b * g.m_items;
      const int q0 = mi * g.r_out;                 // first output row of the item
      const int t_row0 = q0 - g.h_tot;             // position of tile row 0
      const int lin = dummy ? 0 : p.lin;
      const bool edge = t_row0 < 0 || t_row0 + g.mt > lin;     // some tile rows lie outside the utterance
      L2S_RTRACE(0, ntr);
      res_load_x(P, w, slab, x_quad, b, t_row0, sbias + 64, lin);
      L2S_RTRACE(0, ntr);
      publish();
      for (int st = 0; st < g.n_dil; ++st) {
        // ---- phase A: D1 -> S
        L2S_RTRACE(0, ntr);
        mbar_wait(d_full, pd);
        pd ^= 1u;
        tc_fence_after();
        L2S_RTRACE(0, ntr);
        {
          long long* tr = (!CG2 && P.trace && blockIdx.x == 0 && warp == 2 && st == 0 && item < 3 * (int)gridDim.x)
                              ? P.trace + 256 + 16 * (item / (int)gridDim.x) : nullptr;
          if (edge) res_phase<true>(g, w, slab, d1_quad, t_row0, lin, tr);
          else res_phase<false>(g, w, slab, d1_quad, t_row0, lin, tr);
        }
        L2S_RTRACE(0, ntr);
        publish();
        res_prebias_d1(g, w, d1_quad, sbias + (2 * ((st + 1) % g.n_dil)) * 64);   // while c2 runs
        if (st + 1 < g.n_dil) {
          // ---- phase B: X -> S
          mbar_wait(d_full, pd);
          pd ^= 1u;
          tc_fence_after();
          L2S_RTRACE(0, ntr);
          if (edge) res_phase<true>(g, w, slab, x_quad, t_row0, lin);
          else res_phase<false>(g, w, slab, x_quad, t_row0, lin);
          L2S_RTRACE(0, ntr);
          publish();
          res_addbias_x(g, w, x_quad, sbias + (2 * (st + 1) + 1) * 64);           // while the next c1 runs
        }
      }
      // ---- output: X -> global (rows [q0, q0 + r_out) of the tile only)
      const int row_lim = dummy ? 0 : min(p.lin, q0 + g.r_out);
      if (g.c % CW == 0) res_output<CW, MODE, !DUAL && !WIDE>(p, tile, x_quad, b, t_row0, q0, row_lim, g.msub, g.c, w, d_full, pd);
      else res_output<16, MODE, !DUAL && !WIDE>(p, tile, x_quad, b, t_row0, q0, row_lim, g.msub, g.c, w, d_full, pd);
      pd ^= 1u;
      L2S_RTRACE(0, ntr);
      tc_fence_before();
      __syncwarp();
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.span && threadIdx.x == 0) atomicMax(&P.span[1], (unsigned long long)gtime());
  if constexpr (CG2) cluster_sync_all();        // no CTA leaves while its partner may still signal it
  if (warp == 1) { if constexpr (CG2) tmem_dealloc_cg2(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc_dyn(tmem_base, (uint32_t)g.tmem_cols); }
}

// ------------------------------------------------------------------ host side

inline int g_res_iss2 = 0;  // knob res_iss2: the sixteen-warp (WIDE) plans issue their MMAs from TWO warps, each owning half of the tile's
                            // accumulators.  Bit-identical, measured neutral (C = 64, k = 11: 225.5 vs 225.7 us): in this kernel the issuing
                            // warp has its scheduler to itself while the MMAs run, and one warp keeps the tensor pipe fed; off.
inline int g_res_wide = 1;  // knob res_wide: one-CTA-per-SM plans use sixteen epilogue warps (WIDE kernels)
inline int g_res_cg2 = 4;   // whole-ResBlock plans run as CTA pairs issuing cta_group::2 MMAs (knob res_cg2: 0 off, 1 one-CTA-per-SM
                            // plans only, 2 + dual, 3 + quad, 4 + C = 16)

// kind: 0 = one CTA per SM (168 registers, 32-column epilogue chunks), 1 = two (80 registers, <= 112 KB, <= 256 TMEM
// columns), 2 = four CTAs per SM with four epilogue warps each (<= 55 KB, <= 128 TMEM columns): more independent
// tiles in flight for the MMA-light ResBlocks, whose epilogue warps otherwise idle while their own tile's MMAs run.
inline int g_res_skew_iss2 = 1;  // knob res_skew_iss2: skewed plans with two granules and sixteen epilogue warps issue from one warp per granule
inline int g_res_tb = 0;         // knob res_tb: cap on the taps per weight stage of the skewed schedule (0: as res_tc_kernel)
inline int g_res_gmax = 0;       // knob res_gmax: cap on the stages of a head / tail group (0: half the ring)
inline int g_res_ng = 2;         // knob res_ng: granules per tile of the skewed schedule (each with its own barrier pair)
inline int g_res_skew = 0;       // knob res_skew: one-CTA-per-SM plans run the skewed schedule of resq_tc.cuh where its two S slabs fit

inline bool res_plan_with(int c, int k, int n_dil, const int* dil, int lin, int batch, int kind, int msub, ResGeom* out, bool skew = false) {
  const bool dual = kind != 0;
  if (skew && kind != 0) return false;
  ResGeom g{};
  if ((c != 16 && c != 32 && c != 64) || k < 1 || k > kMaxTaps || (k & 1) == 0 || n_dil < 1 || n_dil > kResMaxDil) return false;
  g.c = c; g.k = k; g.n_dil = n_dil;
  g.h2 = (k - 1) / 2;
  int hmax = g.h2;
  for (int s = 0; s < n_dil; ++s) {
    g.dil[s] = dil[s];
    g.h1[s] = dil[s] * (k - 1) / 2;
    g.h_tot += g.h1[s] + g.h2;
    if (g.h1[s] > hmax) hmax = g.h1[s];
  }
  g.pad = (hmax + 7) & ~7;
  g.rb = c * 2;
  g.k16 = g.rb / 32;
  g.lc = c == 16 ? 4 : (c == 32 ? 5 : 6);
  g.dual = dual ? 1 : 0;
  const bool wide = kind == 0 && g_res_wide;
  g.cw = (!dual && c % 32 == 0 && !(skew && wide)) ? 32 : 16;   // skewed + sixteen warps: 16-column output chunks (half the tile memory)
  g.tile_words = 32 * g.cw;
  g.msub = msub;
  g.mt = 128 * msub;
  g.r_out = g.mt - 2 * g.h_tot;
  if (g.r_out < 32) return false;
  int cols = 32;
  while (cols < 2 * msub * c) cols <<= 1;
  if (cols > (kind == 2 ? 128 : (dual ? 256 : 512))) return false;
  g.tmem_cols = cols;
  g.ne = kind == 2 ? 4 : ((kind == 0 && g_res_wide) ? 16 : kTcEpiWarps);
  g.iss2 = (g.ne == 16 && g_res_iss2 && !skew && msub >= 2) ? 1 : 0;   // skewed plans: decided below (an issuer per granule)
  g.ctas_per_sm = kind == 2 ? 4 : (dual ? 2 : 1);
  int tb = 1;
  while (tb < k && tb < 16 && (tb * 2) * c * g.rb <= 16384) tb *= 2;
  if (tb > k) tb = k;
  if (skew && g_res_tb > 0 && g_res_tb < tb) tb = g_res_tb;      // finer weight stages: the granule-outer groups hold fewer bytes
  g.tb = tb;
  g.n_tstages = (k + tb - 1) / tb;
  g.m_items = (lin + g.r_out - 1) / g.r_out;
  g.total_items = batch * g.m_items;
  // CTA pairs issuing cta_group::2 MMAs keep half of every weight stage.  Measured on cfg2 (with default-semantics
  // barrier operations, see mbar_arrive_cluster): C=64 k=11 281 -> 252 us, k=7 185 -> 171, C=32 k=11 185 -> 171, C=16 k=11
  // 146 -> 140; nothing gets slower.
  g.cg2 = (g_res_cg2 && (kind == 0 || g_res_cg2 >= 2 + (kind == 2 ? 1 : 0)) && (c >= 32 || g_res_cg2 >= 4) && g.total_items >= 2) ? 1 : 0;   // knob: 2 also dual, 3 also quad, 4 also C = 16
  g.bstage_bytes = tb * (g.cg2 ? c / 2 : c) * g.rb;
  g.s_bytes = ((g.mt + 2 * g.pad) * g.rb + 1023) & ~1023;
  if ((msub * c) % 32 != 0) return false;   // the phases walk the TMEM region in 32-column units
  if (skew) {
    g.skew = 1;
    const int gmin = c >= 32 ? 1 : 32 / c;      // a granule is a whole number of 32-column units
    if (msub % gmin != 0) return false;
    int ngr = g_res_ng < 1 ? 1 : (g_res_ng > 8 ? 8 : g_res_ng);
    while ((msub / gmin) % ngr != 0) --ngr;
    g.ng = ngr;
    g.gran = msub / ngr;
    g.iss2 = (g_res_skew_iss2 && ngr >= 2 && ngr % 2 == 0 && g.ne == 16) ? 1 : 0;
  }
  const int fixed = 1024 + (skew ? 576 : 192) + 2 * kResMaxDil * 64 * 4 + g.ne * g.tile_words * 4 + (skew ? 2 : 1) * g.s_bytes;   // slack, barriers, biases, tiles, S
  const int budget = kind == 2 ? 55 * 1024 : (dual ? 112 * 1024 : 220 * 1024);
  int sb = 2;
  if (fixed + sb * g.bstage_bytes > budget) return false;
  while (sb < (skew ? 16 : kTcMaxStagesB) && sb < 2 * g.n_tstages && fixed + (sb + 1) * g.bstage_bytes <= budget &&
         (sb + 1) * g.bstage_bytes <= 64 * 1024)
    ++sb;
  g.sb = sb;
  if (skew) {
    // head group: the leading stages whose taps all lie at or left of the centre (their MMAs on a granule read no S row of
    // the next granule); tail group: the trailing stages; both at most half the ring (the producer keeps loading ahead)
    int gmax = sb / 2 > 1 ? sb / 2 : 1;
    if (g_res_gmax > 0 && gmax > g_res_gmax) gmax = g_res_gmax;
    const int nts = g.n_tstages;
    if (nts == 1) { g.gh = 0; g.gt = 1; }
    else {
      int gh = ((k - 1) / 2 + 1) / tb;
      if (gh > gmax) gh = gmax;
      if (gh > nts - 1) gh = nts - 1;
      if (gh < 1) gh = 1;
      g.gh = gh;
      g.gt = nts - gh < gmax ? nts - gh : gmax;
    }
    g.head_fwd = (g.gh * tb - 1 > (k - 1) / 2) ? 1 : 0;
  }
  g.smem_bytes = fixed + sb * g.bstage_bytes;
  g.idesc = umma_idesc_bf16(128u, (uint32_t)c);
  *out = g;
  return true;
}

// Best-scoring tile of each kind; two co-resident CTAs (MMA / epilogue overlap) beat one bigger tile unless the
// halo eats too much of the smaller tile.  mode: 0 auto, 1 force dual, 2 force single, 3 force quad.
inline int g_res_single_pct = 85;   // planner: score discount (percent) of a one-CTA-per-SM plan (knob res_single_pct)

inline int g_res_skew_pct = 100;    // planner: score weight (percent) of a skewed one-CTA-per-SM plan (knob res_skew_pct)
inline int g_res_quad_pct = 115;    // planner: score weight (percent) of a four-CTAs-per-SM plan, 0 = never (knob res_quad_pct).  Measured on
                                    // cfg2: C=32 k=3 92 -> 84 us, C=16 k=7 117 -> 107 us; 200 (quad everywhere it fits) is slower.

inline bool res_plan(int c, int k, int n_dil, const int* dil, int lin, int batch, int mode, int max_msub, ResGeom* out) {
  ResGeom best{};
  double best_score = 0.0;
  int dual_mt = 0;            // tile of the best two-CTAs-per-SM plan (kind 1 is scored before kind 0)
  for (int kind = 2; kind >= 0; --kind) {
    if ((mode == 1 && kind != 1) || (mode == 2 && kind != 0) || (mode == 3 && kind != 2)) continue;
    if (kind == 2 && mode != 3 && g_res_quad_pct <= 0) continue;
    double kind_best = 0.0;
    for (int msub = max_msub < 8 ? max_msub : 8; msub >= 1; --msub) {
      ResGeom g;
      if (!(kind == 0 && g_res_skew && res_plan_with(c, k, n_dil, dil, lin, batch, kind, msub, &g, true)) &&
          !res_plan_with(c, k, n_dil, dil, lin, batch, kind, msub, &g))
        continue;
      // useful rows per computed row, weighted by how well the kind overlaps MMA and epilogue phases (from the
      // measured sweeps); tiles much longer than the utterance waste the rest.  One CTA per SM is discounted less
      // when two CTAs only fit 256-row tiles (C = 64, k = 7: 171 -> 160 us as one 512-row tile per SM).
      const int covered = g.m_items * g.r_out;
      const double single_w = 0.01 * g_res_single_pct + ((dual_mt > 0 && dual_mt <= 256) ? 0.10 : 0.0);
      const double w = kind == 2 ? 0.01 * (g_res_quad_pct > 0 ? g_res_quad_pct : 100) : (kind == 1 ? 1.0 : (g.skew ? 0.01 * g_res_skew_pct : single_w));
      const double score = (double)g.r_out / g.mt * ((double)lin / covered) * w;
      if (kind == 1 && score > kind_best) { kind_best = score; dual_mt = g.mt; }
      if (score > best_score) { best_score = score; best = g; }
    }
  }
  if (best_score <= 0.0) return false;
  *out = best;
  return true;
}

template <int MODE, bool DUAL, bool CG2 = false, bool WIDE = false>
inline cudaError_t launch_res_mode(const ResParams& P, const ResMaps& maps, int grid, cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(res_tc_kernel<MODE, DUAL, CG2, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(res_tc_kernel<MODE, DUAL, CG2, WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(64 + 32 * P.g.ne + (P.g.iss2 ? 32 : 0)));
  cfg.dynamicSmemBytes = (size_t)P.g.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG2 ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, res_tc_kernel<MODE, DUAL, CG2, WIDE>, maps, P);
}

cudaError_t launch_resq_tc(const ResParams& P, const ResMaps& maps, int grid, int mode, cudaStream_t stream);   // resq_tc.cuh / tu_resq_tc.cu

// Defined in tu_res_tc.cu (the only translation unit that instantiates res_tc_kernel); declared everywhere else.
#ifndef L2S_TU_RES_TC
cudaError_t launch_res_tc(const ResParams& P, const ResMaps& maps, int num_ctas, cudaStream_t stream);
#else
cudaError_t launch_res_tc(const ResParams& P, const ResMaps& maps, int num_ctas, cudaStream_t stream) {
  const ResGeom& g = P.g;
  const ConvParams& c = P.c;
  const int cap = num_ctas * g.ctas_per_sm;
  int grid = g.total_items < cap ? g.total_items : cap;
  if (grid < 1) grid = 1;
  if (g.cg2) {                               // CTA pairs: even grid, one pair per two items at most
    const int pairs_needed = (g.total_items + 1) / 2;
    int pairs = cap / 2 < pairs_needed ? cap / 2 : pairs_needed;
    if (pairs < 1) pairs = 1;
    grid = 2 * pairs;
  }
  const int mode = ((c.acc_in || c.div != 1.0f) ? kEpiAcc : 0) | (c.out_raw ? kEpiRaw : 0) | (c.out_act ? kEpiAct : 0) | (c.res ? kEpiRes : 0);
  if (g.skew) return launch_resq_tc(P, maps, grid, mode, stream);
  switch (mode) {
#define L2S_RMODE(m)                                                                                        \
  case m:                                                                                                   \
    if (g.ne == 16) return g.cg2 ? launch_res_mode<m, false, true, true>(P, maps, grid, stream) : launch_res_mode<m, false, false, true>(P, maps, grid, stream); \
    if (g.cg2) return g.dual ? launch_res_mode<m, true, true>(P, maps, grid, stream) : launch_res_mode<m, false, true>(P, maps, grid, stream); \
    return g.dual ? launch_res_mode<m, true>(P, maps, grid, stream) : launch_res_mode<m, false>(P, maps, grid, stream);
    L2S_RMODE(4) L2S_RMODE(6) L2S_RMODE(8) L2S_RMODE(10) L2S_RMODE(12) L2S_RMODE(14)
    L2S_RMODE(7) L2S_RMODE(11) L2S_RMODE(15)      // + a second branch output (p.res) next to the running sum
#undef L2S_RMODE
    default: return cudaErrorInvalidValue;
  }
}
#endif  // L2S_TU_RES_TC

}  // namespace l2s
