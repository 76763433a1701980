// A whole ResBlock1 (speech-resynthesis/models.py:16-45) in one kernel for the narrow stages (C <= 64), TIME-PACKED:
//
//     for d in dilations:  x = x + c2_d( lrelu( c1_d( lrelu(x) ) ) )
//
// res_tc.cuh maps a conv onto M = time, N = C, one MMA per tap: at N = 16..64 an SS-mode tcgen05.mma costs 39..48
// cycles for 8..32 cycles of tensor work (tools/micro/mma_rate.cu) -- the operand fetch, not the tensor pipe, sets the
// pace.  Here P = 128 / C consecutive time steps are packed into one GEMM row ("block"), so that every MMA is
// M = 128 blocks x N = 128 (P time steps x C output channels) x K = 16 and runs at the tensor rate (64 cycles):
//
//   out block i, slot j (time P i + j)  =  sum over input offsets o in [-h, P-1+h]  in(time P i + o) . Wt_o[(j, co)][ci]
//   Wt_o[(j, co)][ci] = w[co][ci][o - j + h]  when that tap exists, else 0          (block-Toeplitz expansion, h = (k-1)/2)
//
// i.e. P + k - 1 "offset" MMAs of K = C instead of P * k tap MMAs of N = C: 18 instead of 88 for C = 16, k = 11.
// A dilated conv (dilation d) is d independent plain convs on the polyphase components x[n d + r], so the input of a
// dilated c1 is stored PHASE-MAJOR (phase r occupies a contiguous run of blocks) and the same offset MMAs apply; the
// epilogue warps, which rewrite the operand slab between any two convs anyway, do the re-layout for free.
//
//   X  (TMEM, fp32, msub x 128 columns)  the residual stream of the tile in natural block order, minus the c2 biases
//                                        added so far (they are per-channel constants, re-added wherever X is read),
//                                        written once from global memory, then only by the c2 MMAs (accumulate)
//   D1 (TMEM, fp32, msub x 128 columns)  accumulator of c1 (first MMA overwrites), in the block order of c1's input
//   S  (shared, bf16)                    two slabs of 128-byte swizzled rows: slab h holds slots [h P/2, (h+1) P/2) of
//                                        every block; the MMA for offset o reads rows shifted by floor(o / P) of slab
//                                        ((o mod P) / (P/2)) at a 2C-byte column offset -- a row-shifted descriptor
//
// Warp roles, barriers, CTA pairs (cta_group::2: each CTA keeps half of every weight stage) and the output epilogue
// are those of res_tc.cuh.  Results differ from the tap-by-tap kernels in the last fp32 bits (summation order).
#pragma once
#include "pair_tc.cuh"

namespace l2s {

constexpr int kPkMaxDil = 3;
constexpr int kPkMaxBr = 3;        // ResBlocks (kernel-size branches of one MRF stage) fused into one launch
constexpr int kPkPadRows = 8;      // zero rows above / below each slab (row shifts reach +-ceil(h / P) <= 4)

struct PkLay {                     // block order of one conv's operand: d = 1 natural, else phase-major by d
  int d;
  uint32_t magic;                  // ceil(2^32 / d): tau / d == __umulhi(tau, magic) for tau < 2^16
  int span;                        // positions per phase (= blocks per phase * P)
};

struct PkGeom {
  int c, n_dil, n_br;              // n_br ResBlocks run back to back on every tile (branch fusion: x and the branch sum stay L2-hot)
  int k[kPkMaxBr], hc[kPkMaxBr];
  int n_off[kPkMaxBr], n_groups[kPkMaxBr], n_tstages[kPkMaxBr];   // offset MMAs, weight groups (Q offsets each), ring stages per conv
  int dil[kPkMaxDil];
  PkLay lay[kPkMaxDil];
  int P, Q;                        // time steps per block / per 128-byte slab row
  int msub, nr, mt, mt_v;          // accumulators, block rows, time steps of a tile, time steps that every layout maps
  int h_tot, hl;                   // summed halo, rounded up to a multiple of P (tile origin stays block aligned)
  int r_out, m_items, total_items;
  int half_bytes, s_bytes;
  int tb, bstage_bytes, sb;
  int tmem_cols, cw, dual, tile_words, ctas_per_sm, cg2;
  int smem_bytes;
};

constexpr int kPkBiasRows = 2 * kPkMaxDil + 1;   // per branch: b1_s / running sum of b2 per step, then the sum of all c2 biases

struct PkParams {
  ConvParams c;                        // output epilogue of the LAST branch in PACKED terms: lin = L / P block rows, ntot = 128
  const float* x;                      // fp32 residual stream entering the block(s), [B][L][C]
  const float* bias_cols;              // [n_br][kPkBiasRows][128]: b1_s replicated over the P slots / running sum of b2_0..b2_s, ..., sum of all b2
  float* acc_buf;                      // n_br > 1: fp32 running branch sum [B][L][C] (written and re-read by the same thread)
  // The same constants by value: they sit in the constant bank and feed the FADDs through the constant cache (read from
  // shared memory they waited behind the MMA operand traffic: 0.3 us per unit measured).  [br][2 s] = b1_s,
  // [br][2 s + 1] = b2_0 + .. + b2_s, [br][2 kPkMaxDil] = sum of all c2 biases (output); one value per channel.
  float bias_ch[kPkMaxBr][kPkBiasRows][64];
  int lin;                             // L (time steps)
  PkGeom g;
  unsigned long long* span;            // debug: [0] min CTA start, [1] max CTA end (globaltimer)
  long long* trace;                    // debug: globaltimer stamps of CTA 0: [0..127] epilogue warp 2, [128..255] MMA warp
};
#define L2S_PTRACE(base, n)                                                          \
  do {                                                                               \
    if (P.trace && blockIdx.x == 0 && lane == 0 && (n) < 128) P.trace[(base) + (n)++] = gtime(); \
  } while (0)

struct PkMaps {
  CUtensorMap w[kPkMaxBr * 2 * kPkMaxDil];   // per branch: Toeplitz weights of c1_0, c2_0, c1_1, c2_1, ...: [n_groups][128][64] bf16
};

struct PkLane { int quad, half, lane; };

// natural time step -> position in the layout (< 0: the layout does not hold it)
__device__ __forceinline__ int pk_pos(const PkLay& L, int tau) {
  if (L.d == 1) return tau;
  const int p = (int)__umulhi((uint32_t)tau, L.magic);
  const int r = tau - p * L.d;
  return p < L.span ? r * L.span + p : -1;
}
// position in the layout -> natural time step (< 0: an unused block of the layout)
__device__ __forceinline__ int pk_tau(const PkLay& L, int x) {
  if (L.d == 1) return x;
  int r = 0;
  while (x >= L.span && r < L.d) { x -= L.span; ++r; }      // d <= 16 phases
  return r < L.d ? x * L.d + r : -1;
}

// One 32-column unit of one block row (values already final: bias added) -> leaky-ReLU(0.1) -> bf16 -> S.
// src_nat: the row index counts blocks in natural order (X, or the D1 of an undilated c1) and the destination is
// layout L; else the row counts blocks of layout L (the D1 of a dilated c1) and the destination is natural order.
template <int C, bool EDGE>
__device__ __forceinline__ void pk_store_unit(uint8_t* slab, int half_bytes, const uint32_t (&r)[32], int row, int c0, bool src_nat,
                                              const PkLay& L, int t_row0, int lin) {
  constexpr int P = 128 / C, Q = P / 2;
  constexpr int LC = C == 16 ? 4 : (C == 32 ? 5 : 6), LP = 7 - LC, LQ = LP - 1;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(0.1f);   // LRELU_SLOPE, models.py:13,38
#pragma unroll
  for (int e = 0; e < 4; ++e) {                               // 8 columns = one 16-byte slot
    const int col = c0 + 8 * e;
    const int j = col >> LC, ch = col & (C - 1);
    const int xs = row * P + j;
    int tau, xd;
    if (src_nat) { tau = xs; xd = pk_pos(L, tau); }
    else { tau = pk_tau(L, xs); xd = tau; }
    if (tau < 0 || xd < 0) continue;
    uint4 pk;
    pk.x = lrelu_bf16x2(__uint_as_float(r[8 * e + 0]), __uint_as_float(r[8 * e + 1]), slope2);
    pk.y = lrelu_bf16x2(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]), slope2);
    pk.z = lrelu_bf16x2(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]), slope2);
    pk.w = lrelu_bf16x2(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]), slope2);
    if (EDGE) {
      const int t = t_row0 + tau;
      if (t < 0 || t >= lin) pk = make_uint4(0u, 0u, 0u, 0u); // the conv's zero padding
    }
    const int drow = xd >> LP, de = xd & (P - 1);
    const int h = de >> LQ, sl = de & (Q - 1);
    *reinterpret_cast<uint4*>(slab + (size_t)h * half_bytes + (size_t)(kPkPadRows + drow) * 128 +
                              (((((sl * C + ch) >> 3)) ^ (drow & 7)) << 4)) = pk;
  }
}

// Phase A / B: TMEM (D1 or X) + per-column constants -> S.
template <int C, bool EDGE>
__device__ __forceinline__ void pk_phase(const PkGeom& g, const PkLane& w, uint8_t* slab, uint32_t t_quad, const float* bias_cols,
                                         bool src_nat, const PkLay& L, int t_row0, int lin) {
  const int n_units = 4 * g.msub;
  for (int u = w.half; u < n_units; u += 2) {
    uint32_t r[32];
    tmem_ld32(t_quad + (uint32_t)(32 * u), r);
    tmem_ld_wait();
    const int c0 = (u & 3) * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bq = *reinterpret_cast<const float4*>(bias_cols + c0 + 4 * j);
      r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + bq.x);
      r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + bq.y);
      r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + bq.z);
      r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + bq.w);
    }
    pk_store_unit<C, EDGE>(slab, g.half_bytes, r, (u >> 2) * 128 + w.quad * 32 + w.lane, c0, src_nat, L, t_row0, lin);
  }
}

// ---- table-driven fast path ---------------------------------------------------------------------------------
// Where a thread's (unit, slot) lands in S depends only on the layout, not on the item, so every thread computes the
// shared-memory offsets of its slots ONCE per kernel (tab[mapping][slot], 16-bit units of 16 bytes, 0xFFFF = the
// layout does not hold this slot) and a phase is TMEM load -> + constants -> leaky-ReLU -> bf16 -> four stores at
// tab ^ piece.  Mappings: 0 load (natural -> layout 0), 1 + 2s phase A of step s (layout s -> natural),
// 2 + 2s phase B of step s (natural -> layout s + 1).  Tiles that touch an utterance edge take the generic path
// above (they also zero what lies outside [0, L)).
constexpr int kPkMaps = 2 * 3;     // the table path covers n_dil <= 3

template <int C>
__device__ __forceinline__ uint32_t pk_slot_offset(const PkGeom& g, int row, int col, bool src_nat, const PkLay& L) {
  constexpr int P = 128 / C, Q = P / 2;
  constexpr int LC = C == 16 ? 4 : (C == 32 ? 5 : 6), LP = 7 - LC, LQ = LP - 1;
  const int j = col >> LC, ch = col & (C - 1);
  const int xs = row * P + j;
  int tau, xd;
  if (src_nat) { tau = xs; xd = pk_pos(L, tau); }
  else { tau = pk_tau(L, xs); xd = tau; }
  if (tau < 0 || xd < 0) return 0xFFFFu;
  const int drow = xd >> LP, de = xd & (P - 1);
  const int h = de >> LQ, sl = de & (Q - 1);
  return (uint32_t)(h * g.half_bytes + (kPkPadRows + drow) * 128 + (((((sl * C + ch) >> 3)) ^ (drow & 7)) << 4)) >> 4;
}

template <int C, int MSUB>
struct PkTab {
  static constexpr int SPU = C == 16 ? 2 : 1;              // slots (time steps, or half of one for C = 64) per 32-column unit
  static constexpr int SLOTS = 2 * MSUB * SPU;             // per thread and phase
  uint32_t v[kPkMaps][SLOTS / 2];                          // two 16-bit offsets per word
  __device__ __forceinline__ uint32_t get(int m, int i) const { return (v[m][i >> 1] >> (16 * (i & 1))) & 0xFFFFu; }
};

template <int C, int MSUB>
__device__ __forceinline__ void pk_build_tab(const PkGeom& g, const PkLane& w, PkTab<C, MSUB>& t) {
  using T = PkTab<C, MSUB>;
  const PkLay nat{1, 0u, g.mt};
#pragma unroll
  for (int m = 0; m < kPkMaps; ++m) {
    const int st = m == 0 ? 0 : (m - 1) >> 1;
    const bool to_nat = m != 0 && ((m - 1) & 1) == 0;       // phase A: layout st -> natural
    const int li = m == 0 ? 0 : (to_nat ? st : st + 1);
    const bool live = li < g.n_dil;
    const PkLay L = live ? g.lay[li] : nat;
#pragma unroll
    for (int i = 0; i < T::SLOTS; i += 2) {
      uint32_t pair = 0;
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int idx = i + k2, ui = idx / T::SPU, sp = idx % T::SPU;
        const int u = w.half + 2 * ui;
        const int row = (u >> 2) * 128 + w.quad * 32 + w.lane;
        const int col = (u & 3) * 32 + sp * (32 / T::SPU);
        const uint32_t off = live ? pk_slot_offset<C>(g, row, col, !to_nat || L.d == 1, L) : 0xFFFFu;
        pair |= off << (16 * k2);
      }
      t.v[m][i >> 1] = pair;
    }
  }
}

__device__ __forceinline__ void ldg256(const float* p, uint32_t* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

// leaky-ReLU -> bf16 -> the unit's four 16-byte pieces at tab ^ piece
template <int C, int MSUB>
__device__ __forceinline__ void pk_store_fast(uint8_t* slab, const uint32_t (&r)[32], const PkTab<C, MSUB>& t, int m, int ui) {
  using T = PkTab<C, MSUB>;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(0.1f);   // LRELU_SLOPE, models.py:13,38
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int sp = T::SPU == 2 ? e >> 1 : 0, pe = T::SPU == 2 ? (e & 1) : e;
    const uint32_t a16 = t.get(m, ui * T::SPU + sp);
    uint4 pk;
    pk.x = lrelu_bf16x2(__uint_as_float(r[8 * e + 0]), __uint_as_float(r[8 * e + 1]), slope2);
    pk.y = lrelu_bf16x2(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]), slope2);
    pk.z = lrelu_bf16x2(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]), slope2);
    pk.w = lrelu_bf16x2(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]), slope2);
    if (a16 != 0xFFFFu) *reinterpret_cast<uint4*>(slab + ((size_t)(a16 ^ (uint32_t)pe) << 4)) = pk;
  }
}

// SMEM_BIAS: the per-column constants come from shared memory ([128] floats, two-CTAs-per-SM kernels); else one value per
// channel from the constant bank (kernel parameters), which keeps them off the shared-memory port the MMAs saturate.
template <int C, int MSUB, bool SMEM_BIAS>
__device__ __forceinline__ void pk_phase_fast(const PkLane& w, uint8_t* slab, uint32_t t_quad, const float* bias, const PkTab<C, MSUB>& t, int m) {
#pragma unroll
  for (int ui = 0; ui < 2 * MSUB; ++ui) {
    const int u = w.half + 2 * ui;
    uint32_t r[32];
    tmem_ld32(t_quad + (uint32_t)(32 * u), r);
    tmem_ld_wait();
    if constexpr (SMEM_BIAS) {
      const float* bc = bias + (u & 3) * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bq = *reinterpret_cast<const float4*>(bc + 4 * j);
        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + bq.x);
        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + bq.y);
        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + bq.z);
        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + bq.w);
      }
    } else {
      const float* bc = bias + (((u & 3) * 32) & (C - 1));
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + bc[j & (C >= 32 ? 31 : C - 1)]);
    }
    pk_store_fast<C, MSUB>(slab, r, t, m, ui);
  }
}

// x (global fp32) -> X (TMEM) and S = lrelu(x): every lane reads its own block row with 256-bit loads (each one a whole
// 32-byte sector, so nothing depends on L1 keeping half-used sectors around); all of a thread's loads are in flight
// before the first is consumed.  Blocks outside [0, L) are read as zeros (lrelu(0) = 0: the conv's zero padding).
template <int C, int MSUB>
__device__ __forceinline__ void pk_load_x_fast(const PkParams& P_, const PkLane& w, uint8_t* slab, uint32_t x_quad, int b, int t_row0, int lin,
                                               const PkTab<C, MSUB>& t) {
  constexpr int P = 128 / C;
  const float* xb = P_.x + (long long)b * lin * C;
  {
#pragma unroll
    for (int ui = 0; ui < 2 * MSUB; ++ui) {
      const int u = w.half + 2 * ui;
      const int tt = t_row0 + ((u >> 2) * 128 + w.quad * 32 + w.lane) * P;
      const bool ok = tt >= 0 && tt < lin;
      const float* src = xb + (long long)tt * C + (u & 3) * 32;
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ok) ldg256(src + 8 * q, &r[8 * q]);
        else {
#pragma unroll
          for (int z = 0; z < 8; ++z) r[8 * q + z] = 0u;
        }
      }
      tmem_st32(x_quad + (uint32_t)(32 * u), r);
      pk_store_fast<C, MSUB>(slab, r, t, 0, ui);
    }
  }
  tmem_st_wait();
}

// x (global fp32) -> X (TMEM) and S = lrelu(x) in the layout of the first c1.  A block row is P * C = 128 consecutive
// floats of the utterance; every lane reads its own row (32 columns = one full 128-byte line per unit).
template <int C>
__device__ __forceinline__ void pk_load_x(const PkParams& P_, const PkLane& w, uint8_t* slab, uint32_t x_quad, int b, int t_row0,
                                          int lin) {
  const PkGeom& g = P_.g;
  constexpr int P = 128 / C;
  const int n_units = 4 * g.msub;
  const float* xb = P_.x + (long long)b * lin * C;
  for (int u = w.half + 2; u < n_units; u += 2) {            // later units: lines into L1 first (no registers held)
    const int t = t_row0 + ((u >> 2) * 128 + w.quad * 32 + w.lane) * P;
    if (t >= 0 && t < lin) prefetch_l1(xb + (long long)t * C + (u & 3) * 32);
  }
  for (int u = w.half; u < n_units; u += 2) {
    const int row = (u >> 2) * 128 + w.quad * 32 + w.lane;
    const int c0 = (u & 3) * 32;
    const int t = t_row0 + row * P;                           // lin and t_row0 are multiples of P: a block is inside or outside
    uint32_t r[32];
    if (t >= 0 && t < lin) {
      const float4* src = reinterpret_cast<const float4*>(xb + (long long)t * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldg(src + j);
        r[4 * j] = __float_as_uint(q.x); r[4 * j + 1] = __float_as_uint(q.y);
        r[4 * j + 2] = __float_as_uint(q.z); r[4 * j + 3] = __float_as_uint(q.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
    tmem_st32(x_quad + (uint32_t)(32 * u), r);
    pk_store_unit<C, false>(slab, g.half_bytes, r, row, c0, true, g.lay[0], t_row0, lin);   // blocks outside [0, L) were read as zeros
  }
  tmem_st_wait();
}

// Output of the block: X (+ the summed c2 biases, p.bias) -> transposed epilogue of conv_tc.cuh in packed terms
// (rows = blocks of the utterance, 128 columns).  Chunks without an output row (the halo at both tile ends) are skipped.
template <int CW, int MODE, bool PIPE>
__device__ __forceinline__ void pk_output(const ConvParams& p, float* tile, uint32_t t_base, int b, int row0, int row_lo, int row_lim,
                                          int msub, const PkLane& w, uint64_t* bar, uint32_t parity) {
  constexpr int LPR = CW / 4;
  constexpr uint32_t kAll = (1u << (CW / 4)) - 1u;
  constexpr int CPS_SH = CW == 32 ? 2 : 3;                    // chunks per 128-column accumulator: 4 or 8
  const int crow = w.lane / LPR, c4 = w.lane % LPR;
  float4* tile4 = reinterpret_cast<float4*>(tile);
  const int n_chunks = msub << CPS_SH;
  // ownership follows the 32-column units of the phases: unit u = half, half + 2, ... is chunk u (CW = 32) or chunks 2u, 2u + 1
  const int first = CW == 32 ? w.half : 2 * w.half;
  auto next_after = [&](int idx) { return CW == 32 ? idx + 2 : ((idx & 1) ? idx + 3 : idx + 1); };
  auto skip = [&](int idx) {
    while (idx < n_chunks) {
      const int r0 = row0 + (idx >> CPS_SH) * 128 + w.quad * 32;
      if (r0 < row_lim && r0 + 32 > row_lo) break;
      idx = next_after(idx);
    }
    return idx;
  };
  auto locate = [&](int idx) {
    const int s_ = idx >> CPS_SH, cc_ = idx & ((1 << CPS_SH) - 1);
    return epi_locate<CW>(p, t_base + (uint32_t)(s_ * 128 + cc_ * CW), b, row0 + s_ * 128 + w.quad * 32, cc_ * CW, crow, c4, row_lim, row_lo);
  };
  const float4 no_res[CW / 4] = {};
  auto finish = [&](const EpiChunk& ch, const float4 (&av)[CW / 4]) {
    epi_stage<CW>(tile4, ch.taddr, w.lane);
    __syncwarp();
    if (__all_sync(0xffffffffu, ch.okmask == kAll)) epi_finish<CW, MODE, true>(p, ch, tile4, no_res, av, crow, c4);
    else epi_finish<CW, MODE, false>(p, ch, tile4, no_res, av, crow, c4);
    __syncwarp();
  };
  int idx = skip(first);
  if constexpr (PIPE) {
    EpiChunk ca{}, cb{};
    float4 ava[CW / 4], avb[CW / 4];
    if (idx < n_chunks) { ca = locate(idx); epi_load_acc<CW, MODE>(p, ca, ava); }
    mbar_wait(bar, parity);
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
    constexpr int RPI = 32 / LPR, ITERS = 32 / RPI;
    auto prefetch_acc = [&](int i2) {
      if constexpr ((MODE & kEpiAcc) != 0) {
        if (p.acc_in && i2 < n_chunks && c4 == 0) {
          const int s_ = i2 >> CPS_SH, cc_ = i2 & ((1 << CPS_SH) - 1);
          const int qa = row0 + s_ * 128 + w.quad * 32 + crow;
#pragma unroll
          for (int i = 0; i < ITERS; ++i) {
            const int q = qa + i * RPI;
            if (q >= row_lo && q < row_lim) prefetch_l1(p.acc_in + ((long long)b * p.lin + q) * p.ntot + cc_ * CW);
          }
        }
      }
    };
    prefetch_acc(idx);
    mbar_wait(bar, parity);
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

__device__ __forceinline__ void ldg256c(const float* p, uint32_t* v) {   // coherent: the running branch sum was written by this thread
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// Output of a branch without the shared-memory transpose: every lane owns one block row (512 contiguous bytes of the
// utterance) and moves it in whole 32-byte sectors (256-bit accesses), 16 columns at a time; the branch-sum sectors of
// the next piece are pulled into L1 while the current one is finished.  bias: one value per channel (constant bank).
template <int C, int MODE>
__device__ __forceinline__ void pk_output_direct(const ConvParams& p, const float* bias, uint32_t t_base, int b, int row0, int row_lo, int row_lim,
                                                 int msub, const PkLane& w, uint64_t* bar, uint32_t parity) {
  const int n_units = 4 * msub;
  const float inv_div = 1.0f / p.div;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(p.slope);
  auto unit_live = [&](int u) {
    const int r0 = row0 + (u >> 2) * 128 + w.quad * 32;
    return r0 < row_lim && r0 + 32 > row_lo;
  };
  auto elem0 = [&](int u, int hu) {   // flat element index of this lane's 16 columns
    const int q = row0 + (u >> 2) * 128 + w.quad * 32 + w.lane;
    return ((long long)b * p.lin + q) * 128 + (u & 3) * 32 + hu * 16;
  };
  auto row_ok = [&](int u) {
    const int q = row0 + (u >> 2) * 128 + w.quad * 32 + w.lane;
    return q >= row_lo && q < row_lim;
  };
  if constexpr ((MODE & kEpiAcc) != 0) {
    for (int u = w.half; u < n_units; u += 2)
      if (p.acc_in && unit_live(u) && row_ok(u)) prefetch_l1(p.acc_in + elem0(u, 0));   // one 128-byte line = the unit's 32 columns
  }
  mbar_wait(bar, parity);
  tc_fence_after();
  for (int u = w.half; u < n_units; u += 2) {
    if (!unit_live(u)) continue;
    const bool ok = row_ok(u);
    const float* bc = bias + (((u & 3) * 32) & (C - 1));
#pragma unroll
    for (int hu = 0; hu < 2; ++hu) {
      uint32_t r[16], a[16];
      const long long e0 = elem0(u, hu);
      if constexpr ((MODE & kEpiAcc) != 0) {
        if (p.acc_in && ok) { ldg256c(p.acc_in + e0, a); ldg256c(p.acc_in + e0 + 8, a + 8); }
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = 0u;
        }
      }
      tmem_ld16(t_base + (uint32_t)(32 * u + 16 * hu), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v = __uint_as_float(r[j]) + bc[(16 * hu + j) & (C >= 32 ? 31 : C - 1)];
        if constexpr ((MODE & kEpiAcc) != 0) v = (v + __uint_as_float(a[j])) * inv_div;
        r[j] = __float_as_uint(v);
      }
      if (ok) {
        if constexpr ((MODE & kEpiRaw) != 0) { stg256(p.out_raw + e0, r); stg256(p.out_raw + e0 + 8, r + 8); }
        if constexpr ((MODE & kEpiAct) != 0) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = lrelu_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), slope2);
          stg256(reinterpret_cast<__nv_bfloat16*>(p.out_act) + e0, pk);
        }
      }
    }
  }
}

// Offset MMAs of one weight stage (groups g0 .. g0 + n_g) for all msub accumulators.
template <int C, bool CG2>
__device__ __forceinline__ void pk_issue_stage(bool leader, int msub, int hc, int n_off, uint32_t desc_hi, uint32_t s_lo, uint32_t half_step,
                                               uint32_t b_lo, int g0, int n_g, uint32_t d_base, bool overwrite) {
  constexpr int P = 128 / C, Q = P / 2, K16 = C / 16;
  constexpr int LP = C == 16 ? 3 : (C == 32 ? 2 : 1), LQ = LP - 1;
  constexpr uint32_t kSlot = (2u * C) >> 4;                   // one time step inside a 128-byte row
  constexpr uint32_t kGroup = ((CG2 ? 64u : 128u) * 128u) >> 4;
  constexpr uint32_t kSub = (128u * 128u) >> 4;
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (((CG2 ? 256u : 128u) >> 4) << 24);
  for (int gi = 0; gi < n_g; ++gi, b_lo += kGroup) {
#pragma unroll
    for (int s = 0; s < Q; ++s) {
      const int oi = (g0 + gi) * Q + s;
      if (oi >= n_off) break;
      const int o = oi - hc;
      const int shift = o >> LP, e = o & (P - 1);
      const uint32_t a_lo = s_lo + (uint32_t)(e >> LQ) * half_step + (uint32_t)((kPkPadRows + shift) * 8) + (uint32_t)(e & (Q - 1)) * kSlot;
      const uint32_t b_s = b_lo + (uint32_t)s * kSlot;
      uint32_t a_sub = a_lo, d_addr = d_base;
      for (int sub = 0; sub < msub; ++sub, a_sub += kSub, d_addr += 128u) {
#pragma unroll
        for (int kk = 0; kk < K16; ++kk) {
          const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_sub + 2u * kk);
          const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_s + 2u * kk);
          const uint32_t acc = (overwrite && oi == 0 && kk == 0) ? 0u : 1u;
          if (leader) {
            if constexpr (CG2) umma_bf16_cg2(d_addr, da, db, kIdesc, acc); else umma_bf16(d_addr, da, db, kIdesc, acc);
          }
        }
      }
    }
  }
}

// All offset MMAs of one conv with every descriptor offset an immediate (HC = (k - 1) / 2 and MSUB compile-time): the
// issuing thread's own instruction stream set the pace of the run-time version above (~250 cycles per MMA measured
// against 64 cycles of tensor work).  Weight stages (TB groups each) are consumed from the ring as they land.
template <int C, int HC, int MSUB, bool CG2, int TB>
__device__ __forceinline__ void pk_issue_conv(bool leader, uint32_t desc_hi, uint32_t desc_lo_fixed, uint32_t s_lo, uint32_t half_step,
                                              uint8_t* stageB, int bstage_bytes, int sb, uint64_t* b_full, uint64_t* b_empty, int& ib,
                                              uint32_t& pb, uint32_t d_base, uint32_t acc_first) {
  constexpr int P = 128 / C, Q = P / 2, K16 = C / 16;
  constexpr int LP = C == 16 ? 3 : (C == 32 ? 2 : 1), LQ = LP - 1;
  constexpr int N_OFF = P + 2 * HC, N_GROUPS = (N_OFF + Q - 1) / Q, N_TS = (N_GROUPS + TB - 1) / TB;
  constexpr uint32_t kSlot = (2u * C) >> 4;
  constexpr uint32_t kGroup = ((CG2 ? 64u : 128u) * 128u) >> 4;
  constexpr uint32_t kSub = (128u * 128u) >> 4;
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (((CG2 ? 256u : 128u) >> 4) << 24);
  const uint32_t s_hi = s_lo + half_step;
#pragma unroll
  for (int ts = 0; ts < N_TS; ++ts) {
    mbar_wait(&b_full[ib], pb);
    tc_fence_after();
    const uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * bstage_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
    for (int gi = 0; gi < TB; ++gi) {
#pragma unroll
      for (int s = 0; s < Q; ++s) {
        constexpr int dummy = 0; (void)dummy;
        const int oi = (ts * TB + gi) * Q + s;               // compile-time after unrolling
        if (oi < N_OFF) {
          const int o = oi - HC;
          const int shift = o >= 0 ? o / P : -((-o + P - 1) / P);
          const int e = o - shift * P;
          const uint32_t a_lo = ((e >> LQ) ? s_hi : s_lo) + (uint32_t)((kPkPadRows + shift) * 8) + (uint32_t)(e & (Q - 1)) * kSlot;
          const uint32_t b_s = b_lo + (uint32_t)gi * kGroup + (uint32_t)s * kSlot;
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
#pragma unroll
            for (int kk = 0; kk < K16; ++kk) {
              const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)sub * kSub + 2u * kk);
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_s + 2u * kk);
              const uint32_t acc = (oi == 0 && kk == 0) ? acc_first : 1u;
              if (leader) {
                if constexpr (CG2) umma_bf16_cg2(d_base + (uint32_t)sub * 128u, da, db, kIdesc, acc);
                else umma_bf16(d_base + (uint32_t)sub * 128u, da, db, kIdesc, acc);
              }
            }
          }
        }
      }
    }
    if (leader) { if constexpr (CG2) umma_commit_cg2(&b_empty[ib], (uint16_t)3); else umma_commit(&b_empty[ib]); }
    if (++ib == sb) { ib = 0; pb ^= 1u; }
  }
}

template <int C, int MODE, bool DUAL, bool CG2>
__global__ void __maxnreg__(DUAL ? 80 : 168)
respk_tc_kernel(const __grid_constant__ PkMaps maps, const __grid_constant__ PkParams P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = P.c;
  const PkGeom& g = P.g;
  constexpr int PP = 128 / C;

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
  float* sbias = reinterpret_cast<float*>(bars + 24);                       // [n_br][2 kPkMaxDil][128]
  float* epi_tiles = sbias + kPkMaxBr * 2 * kPkMaxDil * 128;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (P.span && threadIdx.x == 0) atomicMin(&P.span[0], (unsigned long long)gtime());

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.n_br * 2 * kPkMaxDil; ++i) tma_prefetch_desc(&maps.w[i]);
    for (int i = 0; i < g.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(s_full, (uint32_t)(CG2 ? 2 * kTcEpiWarps : kTcEpiWarps));
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(tmem_slot, (uint32_t)g.tmem_cols); else tmem_alloc_dyn(tmem_slot, (uint32_t)g.tmem_cols); }
  if (warp >= 2) {
    // S starts as zeros: the pad rows stay zero, and blocks a layout never maps must hold finite values (they are
    // multiplied by the zero entries of the Toeplitz weights)
    for (int o = (threadIdx.x - 64) * 16; o < g.s_bytes; o += ((int)blockDim.x - 64) * 16)
      *reinterpret_cast<uint4*>(slab + o) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x - 64; i < g.n_br * 2 * kPkMaxDil * 128; i += (int)blockDim.x - 64) {
      const int br = i / (2 * kPkMaxDil * 128), rest = i - br * (2 * kPkMaxDil * 128);
      sbias[i] = P.bias_cols[br * kPkBiasRows * 128 + rest];
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = g.msub * 128;            // X at [0, acc_cols), D1 at [acc_cols, 2 acc_cols)
  const int crank = CG2 ? (int)cluster_ctarank() : 0;
  const int walkers = CG2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int walk0 = CG2 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int walk_n = CG2 ? (g.total_items + 1) / 2 : g.total_items;
  auto item_of = [&](int wk) { return CG2 ? 2 * wk + crank : wk; };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (Toeplitz weights only)
    const bool leader = elect_one();
    int ib = 0;
    uint32_t pb = 0;
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      const int nxt = wk + walkers < walk_n ? item_of(wk + walkers) : g.total_items;
      if (leader && nxt < g.total_items) {       // next item's x / branch-sum rows into L2, a whole item ahead
        const int nb = nxt / g.m_items;
        const int nq = (nxt - nb * g.m_items) * g.r_out;
        const int lo = max(nq - g.hl, 0), hi = min(nq - g.hl + g.mt, P.lin);
        const long long e0 = ((long long)nb * P.lin + lo) * C;
        const uint32_t bytes = (uint32_t)((hi - lo) * C * 4);
        for (uint32_t o = 0; o < bytes; o += 16384u)
          bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(P.x + e0) + o, min(16384u, bytes - o));
        if ((MODE & kEpiAcc) != 0 && p.acc_in && g.n_br == 1) {
          const int olo = max(nq, 0), ohi = min(nq + g.r_out, P.lin);
          const long long a0 = ((long long)nb * P.lin + olo) * C;
          const uint32_t ab = (uint32_t)((ohi - olo) * C * 4);
          for (uint32_t o = 0; o < ab; o += 16384u)
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.acc_in + a0) + o, min(16384u, ab - o));
        }
      }
      for (int br = 0; br < g.n_br; ++br)
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        const CUtensorMap* wm = &maps.w[br * 2 * kPkMaxDil + cv];
        for (int ts = 0; ts < g.n_tstages[br]; ++ts) {
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if (leader) {
            if constexpr (CG2) {
              if (crank == 0) mbar_expect_tx(&b_full[ib], 2u * (uint32_t)g.bstage_bytes);
              tma_load_3d_cg2(stageB + (size_t)ib * g.bstage_bytes, wm, &b_full[ib], 0, crank * 64, ts * g.tb);
            } else {
              mbar_expect_tx(&b_full[ib], (uint32_t)g.bstage_bytes);
              tma_load_3d(stageB + (size_t)ib * g.bstage_bytes, wm, &b_full[ib], 0, 0, ts * g.tb);
            }
          }
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template(128u);
    const uint32_t desc_hi = (uint32_t)(tmpl >> 32);
    const uint32_t desc_lo_fixed = (uint32_t)tmpl;
    const uint32_t s_lo = desc_lo_fixed | ((smem_u32(slab) & 0x3FFFFu) >> 4);
    const uint32_t half_step = (uint32_t)g.half_bytes >> 4;
    int ib = 0;
    uint32_t pb = 0, ps = 0;
    int ntr = 0;
    auto commit = [&](uint64_t* bar) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); };
    for (int wk = (CG2 && crank != 0) ? walk_n : walk0; wk < walk_n; wk += walkers) {   // CTA pair: the leader issues for both
      for (int br = 0; br < g.n_br; ++br)
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        const bool second = (cv & 1) != 0;
        const int hc = g.hc[br];
        L2S_PTRACE(128, ntr);
        mbar_wait(s_full, ps);
        ps ^= 1u;
        tc_fence_after();
        L2S_PTRACE(128, ntr);
        const uint32_t d_base = second ? tmem_base : tmem_base + (uint32_t)acc_cols;   // c2 accumulates onto X, c1 overwrites D1
        constexpr int MS = DUAL ? 1 : 2;
        const uint32_t acc_first = second ? 1u : 0u;
        if (g.tb == 2 && hc == 1)
          pk_issue_conv<C, 1, MS, CG2, 2>(leader, desc_hi, desc_lo_fixed, s_lo, half_step, stageB, g.bstage_bytes, g.sb, b_full, b_empty, ib, pb, d_base, acc_first);
        else if (g.tb == 2 && hc == 3)
          pk_issue_conv<C, 3, MS, CG2, 2>(leader, desc_hi, desc_lo_fixed, s_lo, half_step, stageB, g.bstage_bytes, g.sb, b_full, b_empty, ib, pb, d_base, acc_first);
        else if (g.tb == 2 && hc == 5)
          pk_issue_conv<C, 5, MS, CG2, 2>(leader, desc_hi, desc_lo_fixed, s_lo, half_step, stageB, g.bstage_bytes, g.sb, b_full, b_empty, ib, pb, d_base, acc_first);
        else
          for (int ts = 0; ts < g.n_tstages[br]; ++ts) {
            mbar_wait(&b_full[ib], pb);
            tc_fence_after();
            const uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * g.bstage_bytes) & 0x3FFFFu) >> 4);
            const int n_g = min(g.tb, g.n_groups[br] - ts * g.tb);
            pk_issue_stage<C, CG2>(leader, g.msub, hc, g.n_off[br], desc_hi, s_lo, half_step, b_lo, ts * g.tb, n_g, d_base, !second);
            if (leader) commit(&b_empty[ib]);
            if (++ib == g.sb) { ib = 0; pb ^= 1u; }
          }
        if (leader) commit(d_full);
        L2S_PTRACE(128, ntr);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    constexpr int CW = DUAL ? 16 : 32;
    PkLane w;
    w.quad = warp & 3;
    w.half = (warp - 2) >> 2;
    w.lane = lane;
    float* tile = epi_tiles + (size_t)(warp - 2) * g.tile_words;
    const uint32_t x_quad = tmem_base + ((uint32_t)(w.quad * 32) << 16);
    const uint32_t d1_quad = x_quad + (uint32_t)acc_cols;
    uint32_t pd = 0;
    int ntr = warp == 2 ? 0 : 128;
    auto publish = [&]() {               // S / X written by this warp are visible to the tensor core
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(s_full, 0u, (uint32_t)crank); else mbar_arrive(s_full); }
    };
    const PkLay nat{1, 0u, g.mt};
    constexpr int MS = DUAL ? 1 : 2;
    PkTab<C, MS> tab;
    const bool use_tab = g.n_dil <= 3;
    if (use_tab) pk_build_tab<C, MS>(g, w, tab);
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      const int item = item_of(wk);
      const bool dummy = item >= g.total_items;    // odd item count: the pair's last partner computes on zeros and stores nothing
      const int b = dummy ? 0 : item / g.m_items;
      const int mi = dummy ? 0 : item - b * g.m_items;
      const int q0 = mi * g.r_out;                 // first output time step of the item (multiple of P)
      const int t_row0 = q0 - g.hl;                // time step of tile position 0 (multiple of P)
      const int lin = dummy ? 0 : P.lin;
      const bool edge = t_row0 < 0 || t_row0 + g.mt > lin;
      const int row_lim = dummy ? 0 : min(P.lin, q0 + g.r_out) / PP;
      for (int br = 0; br < g.n_br; ++br) {
      const float* sb_br = sbias + br * (2 * kPkMaxDil * 128);
      L2S_PTRACE(0, ntr);
      if (use_tab) pk_load_x_fast<C, MS>(P, w, slab, x_quad, b, t_row0, lin, tab);
      else pk_load_x<C>(P, w, slab, x_quad, b, t_row0, lin);
      L2S_PTRACE(0, ntr);
      publish();
      for (int st = 0; st < g.n_dil; ++st) {
        // ---- phase A: D1 (block order of c1's input) + b1 -> S in natural order
        L2S_PTRACE(0, ntr);
        mbar_wait(d_full, pd);
        pd ^= 1u;
        tc_fence_after();
        L2S_PTRACE(0, ntr);
        if (use_tab && !edge) {
          if (st == 0) pk_phase_fast<C, MS, DUAL>(w, slab, d1_quad, DUAL ? sb_br : P.bias_ch[br][0], tab, 1);
          else if (st == 1) pk_phase_fast<C, MS, DUAL>(w, slab, d1_quad, DUAL ? sb_br + 2 * 128 : P.bias_ch[br][2], tab, 3);
          else pk_phase_fast<C, MS, DUAL>(w, slab, d1_quad, DUAL ? sb_br + 4 * 128 : P.bias_ch[br][4], tab, 5);
        } else if (edge) pk_phase<C, true>(g, w, slab, d1_quad, sb_br + (2 * st) * 128, g.lay[st].d == 1, g.lay[st].d == 1 ? nat : g.lay[st], t_row0, lin);
        else pk_phase<C, false>(g, w, slab, d1_quad, sb_br + (2 * st) * 128, g.lay[st].d == 1, g.lay[st].d == 1 ? nat : g.lay[st], t_row0, lin);
        L2S_PTRACE(0, ntr);
        publish();
        if (st + 1 < g.n_dil) {
          // ---- phase B: X + (b2_0 + .. + b2_st) -> S in the block order of the next c1
          mbar_wait(d_full, pd);
          pd ^= 1u;
          tc_fence_after();
          L2S_PTRACE(0, ntr);
          if (use_tab && !edge) {
            if (st == 0) pk_phase_fast<C, MS, DUAL>(w, slab, x_quad, DUAL ? sb_br + 1 * 128 : P.bias_ch[br][1], tab, 2);
            else pk_phase_fast<C, MS, DUAL>(w, slab, x_quad, DUAL ? sb_br + 3 * 128 : P.bias_ch[br][3], tab, 4);
          } else if (edge) pk_phase<C, true>(g, w, slab, x_quad, sb_br + (2 * st + 1) * 128, true, g.lay[st + 1], t_row0, lin);
          else pk_phase<C, false>(g, w, slab, x_quad, sb_br + (2 * st + 1) * 128, true, g.lay[st + 1], t_row0, lin);
          L2S_PTRACE(0, ntr);
          publish();
        }
      }
      // ---- output: X -> global (time steps [q0, q0 + r_out) of the tile only), in block rows of the utterance
      if (br + 1 == g.n_br) {
        if constexpr (DUAL) pk_output<CW, MODE, false>(p, tile, x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
        else pk_output_direct<C, MODE>(p, P.bias_ch[br][2 * kPkMaxDil], x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
      } else {
        // a branch that is not the last: its result goes into (br == 0: starts) the fp32 running sum, which this same
        // thread reads back when the next branch of the tile finishes
        ConvParams pm = p;
        pm.bias = P.bias_cols + (size_t)(br * kPkBiasRows + 2 * kPkMaxDil) * 128;
        pm.out_raw = P.acc_buf;
        pm.out_act = nullptr;
        pm.acc_in = br ? P.acc_buf : nullptr;
        pm.div = 1.0f;
        if constexpr (DUAL) {
          if (br == 0) pk_output<CW, kEpiRaw, false>(pm, tile, x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
          else pk_output<CW, kEpiAcc | kEpiRaw, false>(pm, tile, x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
        } else {
          if (br == 0) pk_output_direct<C, kEpiRaw>(pm, P.bias_ch[br][2 * kPkMaxDil], x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
          else pk_output_direct<C, kEpiAcc | kEpiRaw>(pm, P.bias_ch[br][2 * kPkMaxDil], x_quad, b, t_row0 / PP, q0 / PP, row_lim, g.msub, w, d_full, pd);
        }
      }
      pd ^= 1u;
      L2S_PTRACE(0, ntr);
      tc_fence_before();
      __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.span && threadIdx.x == 0) atomicMax(&P.span[1], (unsigned long long)gtime());
  if constexpr (CG2) cluster_sync_all();
  if (warp == 1) { if constexpr (CG2) tmem_dealloc_cg2(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc_dyn(tmem_base, (uint32_t)g.tmem_cols); }
}

// ------------------------------------------------------------------ host side

inline int g_pk_on = 1;          // knob pack: 0 = never use the time-packed kernel (the tap-by-tap whole-ResBlock kernel runs instead)
inline int g_pk_mode = 0;        // knob pk_mode: 0 auto, 1 two CTAs per SM (msub 1), 2 one CTA per SM (msub 2)
inline int g_pk_cg2 = 1;         // knob pk_cg2: CTA pairs with cta_group::2 MMAs
inline int g_pk_single_pct = 80; // planner: weight (percent) of the one-CTA-per-SM plan (no MMA / epilogue overlap inside one CTA)
inline int g_pk_fuse_br = 1;     // knob pk_fuse: the kernel-size branches of a stage run in ONE launch, back to back on every tile
inline int g_pk_chan_mask = 32;  // knob pk_chan: channel counts (bit mask 16 | 32 | 64) the planner gives to the time-packed kernel.  Measured on
                                 // cfg2 (per stage, us): C = 64 tap-by-tap 507 / packed 522, C = 32 378 / 346, C = 16 306 / 328: at two tiles in
                                 // flight per SM (TMEM) the narrow stages are bound by the MMA <-> epilogue hand-offs, not by MMA time, so the
                                 // 3x cheaper MMAs pay only where a tile carries enough MMA work per hand-off.

// Block-Toeplitz expansion of one Conv1d weight (Cout, Cin, k) (PyTorch layout) into [n_groups][128][64]:
// row n = j * C + co, 128-byte row of group g = Q offsets x C input channels, offset oi = g * Q + s holds
// w[co][ci][oi - j] where 0 <= oi - j < k.
inline void pk_pack_weights(const float* w, int c, int k, std::vector<float>* out) {
  const int P = 128 / c, Q = P / 2;
  const int n_off = P + k - 1, n_groups = (n_off + Q - 1) / Q;
  out->assign((size_t)n_groups * 128 * 64, 0.f);
  for (int oi = 0; oi < n_off; ++oi) {
    const int gq = oi / Q, s = oi % Q;
    for (int j = 0; j < P; ++j) {
      const int tap = oi - j;
      if (tap < 0 || tap >= k) continue;
      for (int co = 0; co < c; ++co)
        for (int ci = 0; ci < c; ++ci)
          (*out)[((size_t)gq * 128 + (j * c + co)) * 64 + s * c + ci] = w[((size_t)co * c + ci) * k + tap];
    }
  }
}

// n_br ResBlocks (kernel sizes ks[]) with the same dilation list run back to back on every tile; the tile geometry is
// that of the widest halo.
inline bool pk_plan_with(int c, int n_br, const int* ks, int n_dil, const int* dil, int lin, int batch, int kind, PkGeom* out) {
  PkGeom g{};
  if ((c != 16 && c != 32 && c != 64) || n_br < 1 || n_br > kPkMaxBr || n_dil < 1 || n_dil > kPkMaxDil) return false;
  g.c = c; g.n_dil = n_dil; g.n_br = n_br;
  g.P = 128 / c; g.Q = g.P / 2;
  if (lin % g.P != 0) return false;                                   // utterance rows must be whole blocks
  int hc_max = 0, max_ts = 0;
  for (int b = 0; b < n_br; ++b) {
    const int k = ks[b];
    if (k < 1 || k > kMaxTaps || (k & 1) == 0) return false;
    g.k[b] = k;
    g.hc[b] = (k - 1) / 2;
    if ((g.hc[b] + g.P - 1) / g.P > kPkPadRows / 2) return false;
    g.n_off[b] = g.P + k - 1;
    g.n_groups[b] = (g.n_off[b] + g.Q - 1) / g.Q;
    if (g.n_groups[b] < 2) return false;
    if (g.hc[b] > hc_max) hc_max = g.hc[b];
  }
  g.tb = 2;
  for (int b = 0; b < n_br; ++b) {
    g.n_tstages[b] = (g.n_groups[b] + g.tb - 1) / g.tb;
    if (g.n_tstages[b] > max_ts) max_ts = g.n_tstages[b];
  }
  g.dual = kind == 1 ? 1 : 0;
  g.msub = g.dual ? 1 : 2;
  g.nr = 128 * g.msub;
  g.mt = g.nr * g.P;
  g.mt_v = g.mt;
  for (int s = 0; s < n_dil; ++s) {
    const int d = dil[s];
    if (d < 1 || d > 16) return false;
    g.dil[s] = d;
    g.h_tot += (d + 1) * hc_max;
    const int nb = g.nr / d;                                          // blocks per phase
    g.lay[s].d = d;
    g.lay[s].magic = d == 1 ? 0u : (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d);
    g.lay[s].span = d == 1 ? g.mt : nb * g.P;
    if (nb < 1) return false;
    if (d * nb * g.P < g.mt_v) g.mt_v = d * nb * g.P;
    for (int t = 0; t < g.mt && d > 1; ++t)                            // the multiply-high division must be exact on the tile
      if ((int)(((uint64_t)(uint32_t)t * g.lay[s].magic) >> 32) != t / d) return false;
  }
  g.hl = (g.h_tot + g.P - 1) / g.P * g.P;
  g.r_out = (g.mt_v - g.h_tot - g.hl) / g.P * g.P;
  if (g.r_out < g.P) return false;
  g.m_items = (lin + g.r_out - 1) / g.r_out;
  g.total_items = batch * g.m_items;
  g.half_bytes = ((g.nr + 2 * kPkPadRows) * 128 + 1023) & ~1023;
  g.s_bytes = 2 * g.half_bytes;
  g.cw = g.dual ? 16 : 32;
  g.tile_words = 32 * g.cw;
  g.ctas_per_sm = g.dual ? 2 : 1;
  g.tmem_cols = 2 * g.msub * 128;
  g.cg2 = (g_pk_cg2 && g.total_items >= 2) ? 1 : 0;
  g.bstage_bytes = g.tb * (g.cg2 ? 64 : 128) * 128;
  const int fixed = 1024 + 192 + kPkMaxBr * 2 * kPkMaxDil * 128 * 4 + kTcEpiWarps * g.tile_words * 4 + g.s_bytes;   // slack, barriers, constants, tiles, S
  const int budget = g.dual ? 113 * 1024 : 220 * 1024;
  int sb = 2;
  if (fixed + sb * g.bstage_bytes > budget) return false;
  while (sb < kTcMaxStagesB && sb < 2 * max_ts && fixed + (sb + 1) * g.bstage_bytes <= budget && (sb + 1) * g.bstage_bytes <= 96 * 1024) ++sb;
  g.sb = sb;
  g.smem_bytes = fixed + sb * g.bstage_bytes;
  *out = g;
  return true;
}

inline bool pk_plan(int c, int n_br, const int* ks, int n_dil, const int* dil, int lin, int batch, PkGeom* out) {
  if (!g_pk_on || !(g_pk_chan_mask & c)) return false;
  PkGeom best{};
  double best_score = 0.0;
  for (int kind = 1; kind >= 0; --kind) {
    if ((g_pk_mode == 1 && kind != 1) || (g_pk_mode == 2 && kind != 0)) continue;
    PkGeom g;
    if (!pk_plan_with(c, n_br, ks, n_dil, dil, lin, batch, kind, &g)) continue;
    const int covered = g.m_items * g.r_out;
    const double score = (double)g.r_out / g.mt * ((double)lin / covered) * (kind == 1 ? 1.0 : 0.01 * g_pk_single_pct);
    if (score > best_score) { best_score = score; best = g; }
  }
  if (best_score <= 0.0) return false;
  *out = best;
  return true;
}

template <int C, int MODE, bool DUAL, bool CG2>
inline cudaError_t launch_respk_mode(const PkParams& P, const PkMaps& maps, int grid, cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(respk_tc_kernel<C, MODE, DUAL, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(respk_tc_kernel<C, MODE, DUAL, CG2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)kTcThreads);
  cfg.dynamicSmemBytes = (size_t)P.g.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG2 ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, respk_tc_kernel<C, MODE, DUAL, CG2>, maps, P);
}

// Output modes the kernel is instantiated for: raw (first branch), acc + raw (middle branches, last stage), acc + act (last
// branch of a stage that feeds an upsampler), acc + raw + act (the same with the debug tap).
inline bool pk_mode_supported(int mode) { return mode == 4 || mode == 6 || mode == 10 || mode == 14; }

// One translation unit per channel count (tu_respk16/32/64.cu) instantiates the kernels of launch_respk_c<C>.
template <int C>
cudaError_t launch_respk_c(const PkParams& P, const PkMaps& maps, int grid, int mode, cudaStream_t stream);

#ifdef L2S_TU_RESPK_C
template <>
cudaError_t launch_respk_c<L2S_TU_RESPK_C>(const PkParams& P, const PkMaps& maps, int grid, int mode, cudaStream_t stream) {
  constexpr int C = L2S_TU_RESPK_C;
  const PkGeom& g = P.g;
  switch (mode) {
#define L2S_KMODE(m)                                                                                                          \
  case m:                                                                                                                     \
    if (g.cg2) return g.dual ? launch_respk_mode<C, m, true, true>(P, maps, grid, stream) : launch_respk_mode<C, m, false, true>(P, maps, grid, stream); \
    return g.dual ? launch_respk_mode<C, m, true, false>(P, maps, grid, stream) : launch_respk_mode<C, m, false, false>(P, maps, grid, stream);
    L2S_KMODE(4) L2S_KMODE(6) L2S_KMODE(10) L2S_KMODE(14)
#undef L2S_KMODE
    default: return cudaErrorInvalidValue;
  }
}
#endif

inline cudaError_t launch_respk_tc(const PkParams& P, const PkMaps& maps, int num_ctas, cudaStream_t stream) {
  const PkGeom& g = P.g;
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
  const int mode = ((c.acc_in || c.div != 1.0f) ? kEpiAcc : 0) | (c.out_raw ? kEpiRaw : 0) | (c.out_act ? kEpiAct : 0);
  if (g.c == 64) return launch_respk_c<64>(P, maps, grid, mode, stream);
  if (g.c == 32) return launch_respk_c<32>(P, maps, grid, mode, stream);
  return launch_respk_c<16>(P, maps, grid, mode, stream);
}

}  // namespace l2s
