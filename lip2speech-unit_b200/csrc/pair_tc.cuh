// Fused ResBlock1 step on tcgen05:   y = x + c2( lrelu( c1( lrelu(x) ) ) )
// (speech-resynthesis/models.py:34-41, one (dilation d, 1) pair of a ResBlock1).
//
// Unfused, c1 writes its activated output to HBM and c2 reads it back; here it goes
// TMEM -> registers (+bias1, leaky-ReLU, bf16) -> SHARED MEMORY in the K-major swizzled
// operand layout -> second tcgen05 GEMM, and only c2's epilogue touches global memory.
//
// Per work item (MT = 128 * msub rows of the intermediate T, R = MT - (k-1) output rows):
//   T[i]   = lrelu(b1 + sum_j W1[j] . xa[t0 + i - h1 + j*d]),  t0 = q0 - h2,  zero outside [0, L)
//   out[r] = b2 + sum_j W2[j] . T[r + j]  + x[q0 + r]  (+ branch sum, mean, lrelu ...)   r < R
// with h1 = d (k-1)/2, h2 = (k-1)/2.  The last k-1 rows of each c2 tile would need T rows that
// this item did not compute; they are simply not stored (items advance by R rows).
//
//   TMA producer (warp 0): xa halo slab per 64-channel chunk; W1 stages, then W2 stages
//   MMA issuer  (warp 1): c1 -> D1 (TMEM cols [0, msub*C)), commit d1_full;
//                         wait t_full (+ d2_empty), c2 from the T slab -> D2, commit d2_full
//   epilogue (warps 2..9): phase 1  D1 -> T slab (fence.proxy.async, arrive t_full)
//                          phase 2  D2 -> global (same transposed epilogue as conv_tc.cuh)
//
// Weight re-streaming is what bounds the C >= 128 layers (every 118..246-row item needs all taps of
// both convs: ~18 TB/s of L2->SM traffic at C = 256 if each CTA fetched its own copy).  With
// g.cluster == 2 the grid is launched as CTA pairs; each CTA TMA-loads HALF of every weight stage
// and multicasts it to both, a stage is released when BOTH CTAs' MMAs have consumed it
// (tcgen05.commit multicast), and the pair walks its items in lockstep (a missing item becomes a
// dummy whose results are not stored).
#pragma once
#include "conv_tc.cuh"

namespace l2s {

inline int g_pair_pref = 0;   // planning experiment switch (knob pair_pref)
inline int g_pair_cg2 = 1;    // cluster plans run cta_group::2 MMAs (knob cg2)

struct PairGeom {
  int c;             // channels (= cin = cout = MMA N)
  int k, dil;        // kernel size, dilation of c1 (c2 has dilation 1)
  int h1, h2;        // halos of c1 and c2
  int rb, kc, k16;
  int msub, mt, r_out;   // accumulators per item, T rows per item, output rows per item
  int m_items;           // ceil(L / r_out)
  int a_rows;            // slab rows = n_loads * box_rows >= mt + 2 h1
  int box_rows, n_loads, slab_bytes, sa;
  int t_rows;            // T slab rows per chunk (mt + 2 h2 rounded up to 8)
  int t_chunk_bytes;     // t_rows * rb
  int tb, n_tstages, bstage_bytes, sb;
  int tmem_cols, cw, total_items;
  int alias_at;          // 1: the A-slab ring lives inside the T-slab region (c1's inputs are dead once T is written)
  int region_bytes;      // shared memory of the A ring + T slab (max of the two when aliased)
  int cluster;           // 1, or 2: CTA pairs fetch each weight stage from L2 once (TMA multicast)
  int cg2;               // cluster == 2 only: the pair runs cta_group::2 MMAs (M = 256: both CTAs' row tiles in one instruction,
                         // each CTA holds and supplies HALF of every weight stage; only the leader CTA issues)
  int dual;              // planned so that two CTAs share one SM (<= 110 KB smem, <= 256 TMEM columns, 80 registers)
  int tile_words;        // fp32 words of one warp's transpose tile (32 rows x cw), or of its TMA staging (epi_tma)
  int epi_tma;           // 1: phase 2 streams the residual in and the results out with TMA through per-warp staging
  uint32_t idesc;
  int smem_bytes;
};

struct PairParams {
  ConvParams c;          // epilogue of c2: bias = b2, res = x, acc_in, out_raw, out_act, div, slope; lin = mrows = L, ntot = C
  const float* bias1;
  PairGeom g;
  unsigned long long* span;   // debug: [0] min CTA start, [1] max CTA end (globaltimer), null in production
  long long* trace;      // debug timestamps of CTA 0 (see conv_tc.cuh), null in production
  // Chained steps of a ResBlock (same k => same item grid): every item of a step counts its finished epilogue warps in
  // done_flags[item]; the NEXT step is launched programmatically (it may occupy SMs as soon as CTAs of this grid exit),
  // does not wait for this grid as a whole, and loads an item's input slab once wait_flags[item - 1 .. item + 1] (the
  // items its halo touches, same utterance) are complete.  Null: plain stream order.
  int* done_flags;
  const int* wait_flags;
};

// swizzled 16-byte slot index inside a T row (matches the TMA / UMMA 128B, 64B, 32B swizzles)
__device__ __forceinline__ int t_swz(int rb, int row, int chunk16) {
  if (rb == 128) return chunk16 ^ (row & 7);
  if (rb == 64) return chunk16 ^ ((row >> 1) & 3);
  return chunk16 ^ ((row >> 2) & 1);
}

// Phase 1 of one [32 rows x CW columns] chunk of D1: + bias1, leaky-ReLU(0.1), bf16, into the T slab.
template <int CW>
__device__ __forceinline__ void pair_phase1_chunk(const PairParams& P, uint8_t* t_slab, uint32_t taddr, int i_row, int c0,
                                                  bool valid) {
  const PairGeom& g = P.g;
  // bias first: its global-load latency hides behind the TMEM load
  float4 bias[CW / 4];
  {
    const float4* b4 = reinterpret_cast<const float4*>(P.bias1 + c0);
#pragma unroll
    for (int s = 0; s < CW / 4; ++s) bias[s] = __ldg(b4 + s);
  }
  uint32_t r[CW];
  if constexpr (CW == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
  tmem_ld_wait();
  const int ch_per_chunk = g.rb >> 1;                       // channels per K chunk
  const int kc2 = c0 / ch_per_chunk;
  const int first16 = ((c0 - kc2 * ch_per_chunk) * 2) >> 4; // first 16-byte slot of these columns in the row
  uint8_t* row_ptr = t_slab + (size_t)kc2 * g.t_chunk_bytes + (size_t)i_row * g.rb;
#pragma unroll
  for (int s = 0; s < CW / 8; ++s) {                        // 8 columns = 16 bytes of bf16
    const float4 ba = bias[2 * s], bb = bias[2 * s + 1];
    float v[8];
    v[0] = __uint_as_float(r[8 * s + 0]) + ba.x; v[1] = __uint_as_float(r[8 * s + 1]) + ba.y;
    v[2] = __uint_as_float(r[8 * s + 2]) + ba.z; v[3] = __uint_as_float(r[8 * s + 3]) + ba.w;
    v[4] = __uint_as_float(r[8 * s + 4]) + bb.x; v[5] = __uint_as_float(r[8 * s + 5]) + bb.y;
    v[6] = __uint_as_float(r[8 * s + 6]) + bb.z; v[7] = __uint_as_float(r[8 * s + 7]) + bb.w;
    uint4 pk;
    uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
    const __nv_bfloat162 slope2 = __float2bfloat162_rn(0.1f);                  // LRELU_SLOPE, models.py:13,38
#pragma unroll
    for (int e = 0; e < 4; ++e) pw[e] = valid ? lrelu_bf16x2(v[2 * e], v[2 * e + 1], slope2) : 0u;
    *reinterpret_cast<uint4*>(row_ptr + (t_swz(g.rb, i_row, first16 + s) << 4)) = pk;
  }
}

// Same arithmetic with the per-chunk overheads removed: bias read from shared memory (broadcast LDS instead of a
// global load per chunk), lane-constant swizzle term, shifts instead of a division, and the zero fill for rows
// outside [0, L) only compiled into the path taken by tiles that touch an utterance edge.
template <int CW, bool EDGE>
__device__ __forceinline__ void pair_phase1_lean(uint8_t* t_slab, int t_chunk_bytes, int rb, int cpc_shift, const float* sbias1,
                                                 uint32_t taddr, int i_row, int c0, int sw, bool valid) {
  uint32_t r[CW];
  if constexpr (CW == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
  tmem_ld_wait();
  uint8_t* row_ptr = t_slab + (size_t)(c0 >> cpc_shift) * t_chunk_bytes + (size_t)i_row * rb;
  const int first16 = (c0 & ((1 << cpc_shift) - 1)) >> 3;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(0.1f);                    // LRELU_SLOPE, models.py:13,38
#pragma unroll
  for (int s = 0; s < CW / 8; ++s) {
    const float4 ba = *reinterpret_cast<const float4*>(sbias1 + c0 + 8 * s);
    const float4 bb = *reinterpret_cast<const float4*>(sbias1 + c0 + 8 * s + 4);
    uint4 pk;
    pk.x = lrelu_bf16x2(__uint_as_float(r[8 * s + 0]) + ba.x, __uint_as_float(r[8 * s + 1]) + ba.y, slope2);
    pk.y = lrelu_bf16x2(__uint_as_float(r[8 * s + 2]) + ba.z, __uint_as_float(r[8 * s + 3]) + ba.w, slope2);
    pk.z = lrelu_bf16x2(__uint_as_float(r[8 * s + 4]) + bb.x, __uint_as_float(r[8 * s + 5]) + bb.y, slope2);
    pk.w = lrelu_bf16x2(__uint_as_float(r[8 * s + 6]) + bb.z, __uint_as_float(r[8 * s + 7]) + bb.w, slope2);
    if (EDGE && !valid) pk = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(row_ptr + (((first16 + s) ^ sw) << 4)) = pk;
  }
}

// ---------------------------------------------------------------------------------------------------
// Asynchronous phase 2 (DUAL plans, modes residual + raw [+ act]).  Per warp and 16-column chunk:
//   residual chunk  [32 rows x 64 B]  global -> staging   by TMA, one chunk ahead (mbarrier res_full[buf])
//   each lane owns one ROW: TMEM row + bias + residual (LDS.128, swizzle-64B slots) -> written back in place,
//   activated bf16 copy into a [32 x 32 B] buffer (swizzle-32B slots)
//   fence.proxy.async, then ONE lane issues the TMA stores (raw fp32, act bf16) as a bulk group.
// No transpose round trip, no LSU global traffic, rows >= r_out of the item are cut by a shorter "tail" box,
// rows >= L by the tensor bounds.
constexpr int kTmaStageBytes = 2 * 2048 + 1024;   // per warp: 2 residual/raw buffers + 1 act buffer

// chunk i of this warp -> (accumulator s, 16-column block cc); cps = 16-column chunks per accumulator (1, 2 or 4)
struct TmaWalk {
  int cps_sh, n_valid, row_q, half;
  __device__ __forceinline__ void at(int i, int& s, int& cc) const {
    const int lin = half + 2 * i;
    s = lin >> cps_sh;
    cc = lin - (s << cps_sh);
  }
};

__device__ __forceinline__ TmaWalk tma_walk(const PairGeom& g, int q0, int row_lim, int quad, int half) {
  TmaWalk w;
  const int cps = g.c >> 4;
  w.cps_sh = cps == 4 ? 2 : (cps == 2 ? 1 : 0);
  w.half = half;
  w.row_q = q0 + quad * 32;
  const int n_all = (g.msub * cps - half + 1) >> 1;          // my chunks: linear index half, half + 2, ...
  w.n_valid = 0;                                             // chunks whose first row is inside the item (a prefix)
  for (int i = 0; i < n_all; ++i) w.n_valid += (w.row_q + (((half + 2 * i) >> w.cps_sh) << 7)) < row_lim ? 1 : 0;
  return w;
}

// Residual prefetch of an item, issued BEFORE phase 1: both staging buffers are filled while c1's epilogue and
// c2's MMAs run, so phase 2 never waits for DRAM.  (With one chunk of look-ahead every warp ate a full loaded
// DRAM latency per chunk: 16 warps x 2 KB in flight per SM = ~2.5 TB/s chip-wide, which is what ncu showed.)
__device__ __forceinline__ void tma_prefetch_res(const TmaWalk& w, const CUtensorMap* tmRes, uint8_t* stage,
                                                 uint64_t* res_full, int b, int lane) {
  if (lane != 0 || w.n_valid == 0) return;
  bulk_wait_read<0>();                                       // the previous item's stores have left the staging buffers
  const int n = w.n_valid < 2 ? w.n_valid : 2;
  for (int i = 0; i < n; ++i) {
    int s, cc;
    w.at(i, s, cc);
    mbar_expect_tx(&res_full[i], 2048u);
    tma_load_3d(stage + i * 2048, tmRes, &res_full[i], cc * 16, w.row_q + s * 128, b);
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_item_tma(const ConvParams& p, const PairGeom& g, const TmaWalk& w,
                                                  const CUtensorMap* tmRes, const CUtensorMap* tmRaw,
                                                  const CUtensorMap* tmRawT, const CUtensorMap* tmAct,
                                                  const CUtensorMap* tmActT, uint8_t* stage, uint64_t* res_full,
                                                  uint32_t (&ph)[2], uint32_t t2, int b, int quad, int lane,
                                                  uint64_t* bar, uint32_t parity, long long* tr = nullptr) {
  // tr (debug, warp 2 of CTA 0): [0] enter, [1] D2 ready, then per chunk i (<2): [2+3i] tmem ld done, [3+3i] residual
  // landed, [4+3i] stores issued
  constexpr bool kAct = (MODE & kEpiAct) != 0;
  if (tr && lane == 0) tr[0] = gtime();
  uint8_t* resb0 = stage;
  uint8_t* resb1 = stage + 2048;
  uint4* actb = reinterpret_cast<uint4*>(stage + 4096);
  mbar_wait(bar, parity);                                    // D2 complete
  tc_fence_after();
  if (tr && lane == 0) tr[1] = gtime();
  const int sw64 = (lane >> 1) & 3, sw32 = (lane >> 2) & 1;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(p.slope);
  for (int i = 0; i < w.n_valid; ++i) {
    int s, cc;
    w.at(i, s, cc);
    const int buf = i & 1;
    float4 bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = __ldg(reinterpret_cast<const float4*>(p.bias + cc * 16) + j);
    uint32_t r[16];
    tmem_ld16(t2 + (uint32_t)(s * g.c + cc * 16), r);
    tmem_ld_wait();
    if (tr && lane == 0 && i < 2) tr[2 + 3 * i] = gtime();
    mbar_wait(&res_full[buf], ph[buf]);
    ph[buf] ^= 1u;
    if (tr && lane == 0 && i < 2) tr[3 + 3 * i] = gtime();
    float4* rb4 = reinterpret_cast<float4*>(buf ? resb1 : resb0);
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int slot = lane * 4 + (j ^ sw64);
      const float4 rs = rb4[slot];
      float4 v;
      v.x = __uint_as_float(r[4 * j + 0]) + bias[j].x + rs.x;
      v.y = __uint_as_float(r[4 * j + 1]) + bias[j].y + rs.y;
      v.z = __uint_as_float(r[4 * j + 2]) + bias[j].z + rs.z;
      v.w = __uint_as_float(r[4 * j + 3]) + bias[j].w + rs.w;
      rb4[slot] = v;                                         // raw result, in place
      if constexpr (kAct) {
        pk[2 * j] = lrelu_bf16x2(v.x, v.y, slope2);
        pk[2 * j + 1] = lrelu_bf16x2(v.z, v.w, slope2);
      }
    }
    // the act buffer is single: the previous chunk's act store must have finished READING it
    if (i > 0) {
      if (lane == 0) {
        bulk_wait_read<0>();
        if (i + 1 < w.n_valid && i + 1 >= 2) {               // geometries with more than two chunks per warp: one ahead
          int s2, cc2;
          w.at(i + 1, s2, cc2);
          mbar_expect_tx(&res_full[(i + 1) & 1], 2048u);
          tma_load_3d(((i + 1) & 1) ? resb1 : resb0, tmRes, &res_full[(i + 1) & 1], cc2 * 16, w.row_q + s2 * 128, b);
        }
      }
      __syncwarp();
    }
    if constexpr (kAct) {
      actb[lane * 2 + (0 ^ sw32)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      actb[lane * 2 + (1 ^ sw32)] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async_smem();                                // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
      const bool tail = s == g.msub - 1 && quad == 3;        // the item's last (k-1) rows are not computable here
      tma_store_3d(tail ? tmRawT : tmRaw, buf ? resb1 : resb0, cc * 16, w.row_q + s * 128, b);
      if constexpr (kAct) tma_store_3d(tail ? tmActT : tmAct, actb, cc * 16, w.row_q + s * 128, b);
      bulk_commit();
      if (tr && i < 2) tr[4 + 3 * i] = gtime();
    }
  }
}

// MMAs of one weight stage for all msub accumulators with every stride and the instruction descriptor as
// immediates (C >= 64: 128-byte operand rows, four K = 16 slices per 64-channel chunk).
template <int C, bool CG2 = false>
__device__ __forceinline__ void pair_issue_stage(bool leader, int msub, uint32_t desc_hi, uint32_t a_lo, uint32_t tap_step,
                                                 uint32_t b_lo, int tap0, int t_end, uint32_t first_or, uint32_t d_base) {
  constexpr uint32_t kSubStep = (128u * 128u) >> 4;
  constexpr uint32_t kTapW = ((uint32_t)(CG2 ? C / 2 : C) * 128u) >> 4;   // CTA pair: a CTA holds half of the weight rows
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)C >> 3) << 17) | (((CG2 ? 256u : 128u) >> 4) << 24);
  uint32_t a_tap = a_lo + (uint32_t)tap0 * tap_step;
  for (int t = 0; t < t_end; ++t, b_lo += kTapW, a_tap += tap_step) {
    const uint32_t first = first_or | (uint32_t)(tap0 + t);
    uint32_t a_sub = a_tap;
    uint32_t d_addr = d_base;
    for (int s = 0; s < msub; ++s, a_sub += kSubStep, d_addr += (uint32_t)C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_sub + 2u * k);
        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
        if (leader) {
          if constexpr (CG2) umma_bf16_cg2(d_addr, da, db, kIdesc, (first | (uint32_t)k) != 0u ? 1u : 0u);
          else umma_bf16(d_addr, da, db, kIdesc, (first | (uint32_t)k) != 0u ? 1u : 0u);
        }
      }
    }
  }
}

template <int MODE, bool DUAL, bool EPI_TMA, bool CG2 = false>
__global__ void __maxnreg__(DUAL ? 80 : 168)
pair_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
               const __grid_constant__ CUtensorMap tmRaw, const __grid_constant__ CUtensorMap tmRawT,
               const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmActT, const PairParams P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = P.c;
  const PairGeom& g = P.g;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slabA = smem;
  uint8_t* slabT = g.alias_at ? smem : slabA + (size_t)g.sa * g.slab_bytes;
  uint8_t* stageB = smem + (size_t)g.region_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageB + (size_t)g.sb * g.bstage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kTcMaxStagesA;
  uint64_t* b_full = a_empty + kTcMaxStagesA;
  uint64_t* b_empty = b_full + kTcMaxStagesB;
  uint64_t* d1_full = b_empty + kTcMaxStagesB;
  uint64_t* t_full = d1_full + 1;
  uint64_t* d2_full = t_full + 1;
  uint64_t* d2_empty = d2_full + 1;
  uint64_t* t_free = d2_empty + 1;      // c2 has finished reading the T slab (gates the aliased A-slab loads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + 1);
  uint64_t* res_full = bars + 40;        // EPI_TMA: 2 residual barriers per epilogue warp
  uint8_t* epi_base = reinterpret_cast<uint8_t*>(bars + 64);
  epi_base += (1024u - (smem_u32(epi_base) & 1023u)) & 1023u;   // swizzle patterns of the staging need aligned bases
  float* sbias1 = reinterpret_cast<float*>(epi_base);           // c1 bias, read by phase 1
  float* epi_tiles = reinterpret_cast<float*>(epi_base + 1024);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (P.span && threadIdx.x == 0) atomicMin(&P.span[0], (unsigned long long)gtime());
  if (P.trace && threadIdx.x == 0 && blockIdx.x < 512) {   // debug: which SM ran this CTA, and when
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    P.trace[768 + blockIdx.x * 3 + 0] = smid;
    P.trace[768 + blockIdx.x * 3 + 1] = gtime();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < g.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < g.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CG2 ? 1u : (uint32_t)g.cluster); }
    mbar_init(d1_full, 1);
    mbar_init(t_full, (CG2 ? 2 : 1) * kTcEpiWarps);     // CTA pair: the leader's barrier collects both CTAs' epilogue warps
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, (CG2 ? 2 : 1) * kTcEpiWarps);
    mbar_init(t_free, 1);
    if (EPI_TMA) {
      for (int i = 0; i < 2 * kTcEpiWarps; ++i) mbar_init(&res_full[i], 1);
      tma_prefetch_desc(&tmRes); tma_prefetch_desc(&tmRaw); tma_prefetch_desc(&tmRawT);
      tma_prefetch_desc(&tmAct); tma_prefetch_desc(&tmActT);
    }
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(tmem_slot, (uint32_t)g.tmem_cols); else tmem_alloc_dyn(tmem_slot, (uint32_t)g.tmem_cols); }
  if (warp >= 2) {
    // rows [mt, t_rows) of every T chunk are read by the last taps of c2 but never written: zero them once
    const int tail_bytes = (g.t_rows - g.mt) * g.rb;
    for (int kc2 = 0; kc2 < g.kc; ++kc2) {
      uint8_t* base = slabT + (size_t)kc2 * g.t_chunk_bytes + (size_t)g.mt * g.rb;
      for (int o = (threadIdx.x - 64) * 16; o < tail_bytes; o += (kTcThreads - 64) * 16)
        *reinterpret_cast<uint4*>(base + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = threadIdx.x - 64; i < g.c; i += kTcThreads - 64) sbias1[i] = P.bias1[i];
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (g.cluster > 1) cluster_sync_all();      // the partner's barriers exist before anything is multicast to them
  tc_fence_after();
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel's
  // tail; nothing below may read or write its data before it has completed.
  pdl_launch_dependents();
  if (!P.wait_flags) pdl_wait_prior_grid();   // chained step: per-item flags instead (the previous step may still be running)
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = g.msub * g.c;          // D1 at [0, acc_cols), D2 at [acc_cols, 2 acc_cols)
  // item walk: a CTA pair advances in lockstep (pair p handles items 2j + rank, j = p, p + pairs, ...)
  const int crank = g.cluster > 1 ? (int)cluster_ctarank() : 0;
  const int walkers = g.cluster > 1 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int walk0 = g.cluster > 1 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int walk_n = g.cluster > 1 ? (g.total_items + 1) / 2 : g.total_items;
  const uint16_t mc_mask = (uint16_t)((1u << g.cluster) - 1u);
  const int half_rows = g.c / g.cluster;      // weight rows (output channels) this CTA fetches per stage

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0, ptf = 0;
    const uint32_t box_bytes = (uint32_t)(g.box_rows * g.rb);
    int it_no = 0;
    for (int w = walk0; w < walk_n; w += walkers, ++it_no) {
      const int item = g.cluster > 1 ? 2 * w + crank : w;   // may be == total_items (dummy): b == batch, TMA zero fills
      const int b = item / g.m_items;
      const int mi = item - b * g.m_items;
      const int row0 = mi * g.r_out - g.h2 - g.h1;          // first xa row of the slab
      if (P.wait_flags && item < g.total_items) {
        // the previous step's items this slab reads from (halo h1 + h2 < r_out: the neighbours at most)
        if (lane == 0) {
          const int lo = mi > 0 ? mi - 1 : 0, hi = mi + 1 < g.m_items ? mi + 1 : g.m_items - 1;
          for (int nb = lo; nb <= hi; ++nb) {
            const int* f = P.wait_flags + b * g.m_items + nb;
            uint32_t spins = 0;
            while (ld_acquire_gpu(f) < kTcEpiWarps) {
              if (++spins > (1u << 24)) __trap();
              __nanosleep(64);
            }
          }
        }
        __syncwarp();
        fence_proxy_async_all();                            // rows written through the generic proxy, read by TMA
      }
      if (g.alias_at && it_no > 0) {                        // the slabs overwrite the T slab of the previous item
        mbar_wait(t_free, ptf);
        ptf ^= 1u;
      }
      for (int kc = 0; kc < g.kc; ++kc) {
        const int ch0 = kc * (g.rb >> 1);
        if (kc == 0) L2S_TRACE(0, it_no, 0);
        mbar_wait(&a_empty[ia], pa ^ 1u);
        if (kc == 0) L2S_TRACE(0, it_no, 1);
        if (leader) {
          uint8_t* dst = slabA + (size_t)ia * g.slab_bytes;
          if constexpr (CG2) {   // both CTAs' slabs complete on the pair leader's barrier
            if (crank == 0) mbar_expect_tx(&a_full[ia], 2u * (uint32_t)g.slab_bytes);
            for (int l = 0; l < g.n_loads; ++l)
              tma_load_3d_cg2(dst + (size_t)l * box_bytes, &tmA, &a_full[ia], ch0, row0 + l * g.box_rows, b);
          } else {
            mbar_expect_tx(&a_full[ia], (uint32_t)g.slab_bytes);
            for (int l = 0; l < g.n_loads; ++l)
              tma_load_3d(dst + (size_t)l * box_bytes, &tmA, &a_full[ia], ch0, row0 + l * g.box_rows, b);
          }
        }
        if (++ia == g.sa) { ia = 0; pa ^= 1u; }
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if constexpr (CG2) {   // this CTA's half of the stage, at the stage base; bytes counted on the leader's barrier
            if (leader) {
              if (crank == 0) mbar_expect_tx(&b_full[ib], 2u * (uint32_t)g.bstage_bytes);   // a stage slot holds this CTA's half
              tma_load_3d_cg2(stageB + (size_t)ib * g.bstage_bytes, &tmW1, &b_full[ib], ch0, crank * half_rows, ts * g.tb);
            }
          } else if (leader) {
            mbar_expect_tx(&b_full[ib], (uint32_t)g.bstage_bytes);
            if (g.cluster > 1)
              tma_load_3d_mc(stageB + (size_t)ib * g.bstage_bytes + (size_t)crank * half_rows * g.rb, &tmW1, &b_full[ib], ch0,
                             crank * half_rows, ts * g.tb, mc_mask);
            else
              tma_load_3d(stageB + (size_t)ib * g.bstage_bytes, &tmW1, &b_full[ib], ch0, 0, ts * g.tb);
          }
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
      for (int kc = 0; kc < g.kc; ++kc) {
        const int ch0 = kc * (g.rb >> 1);
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if constexpr (CG2) {
            if (leader) {
              if (crank == 0) mbar_expect_tx(&b_full[ib], 2u * (uint32_t)g.bstage_bytes);
              tma_load_3d_cg2(stageB + (size_t)ib * g.bstage_bytes, &tmW2, &b_full[ib], ch0, crank * half_rows, ts * g.tb);
            }
          } else if (leader) {
            mbar_expect_tx(&b_full[ib], (uint32_t)g.bstage_bytes);
            if (g.cluster > 1)
              tma_load_3d_mc(stageB + (size_t)ib * g.bstage_bytes + (size_t)crank * half_rows * g.rb, &tmW2, &b_full[ib], ch0,
                             crank * half_rows, ts * g.tb, mc_mask);
            else
              tma_load_3d(stageB + (size_t)ib * g.bstage_bytes, &tmW2, &b_full[ib], ch0, 0, ts * g.tb);
          }
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template((uint32_t)g.rb);
    const uint32_t desc_hi = (uint32_t)(tmpl >> 32);
    const uint32_t desc_lo_fixed = (uint32_t)tmpl;
    const uint32_t sub_step = (uint32_t)(128 * g.rb) >> 4;
    const uint32_t tapw_step = (uint32_t)(g.c * g.rb) >> 4;
    const uint32_t row_step = (uint32_t)g.rb >> 4;            // one row, in descriptor units
    const uint32_t t_lo0 = desc_lo_fixed | ((smem_u32(slabT) & 0x3FFFFu) >> 4);
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0, pt = 0, pd2 = 0;
    int it_no = 0;
    constexpr bool cg2 = CG2;
    auto commit = [&](uint64_t* bar, bool both) {   // both: the partner CTA waits on its own copy of this barrier too
      if constexpr (CG2) umma_commit_cg2(bar, mc_mask);
      else if (both && g.cluster > 1) umma_commit_mc(bar, mc_mask);
      else umma_commit(bar);
    };
    for (int w = (cg2 && crank != 0) ? walk_n : walk0; w < walk_n; w += walkers, ++it_no) {   // CTA pair: the leader issues for both
      // ---- c1: D1 += xa(slab, row shift j*d) . W1[j]
      for (int kc = 0; kc < g.kc; ++kc) {
        if (kc == 0) L2S_TRACE(1, it_no, 0);
        mbar_wait(&a_full[ia], pa);
        if (kc == 0) L2S_TRACE(1, it_no, 1);
        tc_fence_after();
        const uint32_t a_lo = desc_lo_fixed | ((smem_u32(slabA + (size_t)ia * g.slab_bytes) & 0x3FFFFu) >> 4);
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_full[ib], pb);
          tc_fence_after();
          uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * g.bstage_bytes) & 0x3FFFFu) >> 4);
          const int t_end = min(g.tb, g.k - ts * g.tb);
          if constexpr (CG2) {
            if (g.c == 256) pair_issue_stage<256, true>(leader, g.msub, desc_hi, a_lo, (uint32_t)g.dil * row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base);
            else pair_issue_stage<128, true>(leader, g.msub, desc_hi, a_lo, (uint32_t)g.dil * row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base);
          }
          else if (g.c == 256) { pair_issue_stage<256>(leader, g.msub, desc_hi, a_lo, (uint32_t)g.dil * row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base); }
          else if (g.c == 128) { pair_issue_stage<128>(leader, g.msub, desc_hi, a_lo, (uint32_t)g.dil * row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base); }
          else for (int t = 0; t < t_end; ++t, b_lo += tapw_step) {
            const int tap = ts * g.tb + t;
            uint32_t a_sub = a_lo + (uint32_t)(tap * g.dil) * row_step;
            const uint32_t first = (uint32_t)(kc | tap);
            uint32_t d_addr = tmem_base;
            if (g.k16 == 4) {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<4>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            } else if (g.k16 == 2) {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<2>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            } else {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<1>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            }
          }
          if (leader) commit(&b_empty[ib], true);
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
        if (leader) commit(&a_empty[ia], false);
        if (++ia == g.sa) { ia = 0; pa ^= 1u; }
      }
      if (leader) commit(d1_full, false);
      L2S_TRACE(1, it_no, 2);
      // ---- c2: D2 += T(slab, row shift j) . W2[j]   (T written by the epilogue warps)
      mbar_wait(t_full, pt);
      pt ^= 1u;
      mbar_wait(d2_empty, pd2 ^ 1u);
      pd2 ^= 1u;
      L2S_TRACE(1, it_no, 3);
      tc_fence_after();
      for (int kc = 0; kc < g.kc; ++kc) {
        const uint32_t t_lo = t_lo0 + (uint32_t)(kc * g.t_chunk_bytes >> 4);
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_full[ib], pb);
          tc_fence_after();
          uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * g.bstage_bytes) & 0x3FFFFu) >> 4);
          const int t_end = min(g.tb, g.k - ts * g.tb);
          if constexpr (CG2) {
            if (g.c == 256) pair_issue_stage<256, true>(leader, g.msub, desc_hi, t_lo, row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base + (uint32_t)acc_cols);
            else pair_issue_stage<128, true>(leader, g.msub, desc_hi, t_lo, row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base + (uint32_t)acc_cols);
          }
          else if (g.c == 256) { pair_issue_stage<256>(leader, g.msub, desc_hi, t_lo, row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base + (uint32_t)acc_cols); }
          else if (g.c == 128) { pair_issue_stage<128>(leader, g.msub, desc_hi, t_lo, row_step, b_lo, ts * g.tb, t_end, (uint32_t)kc, tmem_base + (uint32_t)acc_cols); }
          else for (int t = 0; t < t_end; ++t, b_lo += tapw_step) {
            const int tap = ts * g.tb + t;
            uint32_t a_sub = t_lo + (uint32_t)tap * row_step;
            const uint32_t first = (uint32_t)(kc | tap);
            uint32_t d_addr = tmem_base + (uint32_t)acc_cols;
            if (g.k16 == 4) {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<4>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            } else if (g.k16 == 2) {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<2>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            } else {
              for (int s = 0; s < g.msub; ++s, a_sub += sub_step, d_addr += (uint32_t)g.c)
                issue_chunk<1>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first);
            }
          }
          if (leader) commit(&b_empty[ib], true);
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
      if (leader) { commit(d2_full, false); commit(t_free, false); }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float* tile = epi_tiles + (size_t)(warp - 2) * g.tile_words;
    uint32_t pd1 = 0, pd2 = 0;
    uint32_t ph_res[2] = {0u, 0u};
    int it_no = 0;
    int unpublished = -1;   // chained steps: the item whose stores this warp has issued but not yet counted in done_flags
    // Counting an item needs a fence that waits for this warp's stores of it; done right after the stores it would expose
    // their latency once per item, so the count of item i is published after phase 1 of item i + 1 (the stores are long
    // out by then; the consumer's CTAs only become resident as CTAs of this grid exit) and at the end for the last item.
    auto publish_item = [&]() {
      if (unpublished >= 0) {
        fence_proxy_async_all();
        __threadfence();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(P.done_flags + unpublished, 1);
        unpublished = -1;
      }
    };
    for (int w = walk0; w < walk_n; w += walkers, ++it_no) {
      const int item = g.cluster > 1 ? 2 * w + crank : w;
      const bool dummy = item >= g.total_items;             // odd item count: the pair's last partner stores nothing
      const int b = item / g.m_items;
      const int mi = item - b * g.m_items;
      const int q0 = mi * g.r_out;
      TmaWalk tw{};
      if constexpr (EPI_TMA) {   // residual of this item: in flight before phase 1 even starts
        tw = tma_walk(g, q0, dummy ? 0 : min(p.lin, q0 + g.r_out), quad, half);
        tma_prefetch_res(tw, &tmRes, reinterpret_cast<uint8_t*>(epi_tiles) + (size_t)(warp - 2) * kTmaStageBytes,
                         res_full + 2 * (warp - 2), b, lane);
      }
      // ---- phase 1: D1 -> T slab
      mbar_wait(d1_full, pd1);
      if (warp == 2) L2S_TRACE(2, it_no, 0);
      pd1 ^= 1u;
      tc_fence_after();
      {
        const uint32_t t1 = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int cps = g.c / g.cw;
        const bool edge = dummy || q0 - g.h2 < 0 || q0 - g.h2 + g.mt > p.lin;   // some T rows lie outside the utterance
        const int sw = g.rb == 128 ? (lane & 7) : (g.rb == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
        const int cpc_shift = g.rb == 128 ? 6 : (g.rb == 64 ? 5 : 4);           // log2(channels per K chunk)
        int s = 0, cc = half;
        while (cc >= cps) { cc -= cps; ++s; }
        while (s < g.msub) {
          const int i_row = s * 128 + quad * 32 + lane;
          const int t = q0 - g.h2 + i_row;
          const bool valid = t >= 0 && t < p.lin && !dummy;
          if (!DUAL && g.cw == 32) {
            const uint32_t ta = t1 + (uint32_t)(s * g.c + cc * 32);
            if (edge) pair_phase1_lean<32, true>(slabT, g.t_chunk_bytes, g.rb, cpc_shift, sbias1, ta, i_row, cc * 32, sw, valid);
            else pair_phase1_lean<32, false>(slabT, g.t_chunk_bytes, g.rb, cpc_shift, sbias1, ta, i_row, cc * 32, sw, true);
          } else {
            const uint32_t ta = t1 + (uint32_t)(s * g.c + cc * 16);
            if (edge) pair_phase1_lean<16, true>(slabT, g.t_chunk_bytes, g.rb, cpc_shift, sbias1, ta, i_row, cc * 16, sw, valid);
            else pair_phase1_lean<16, false>(slabT, g.t_chunk_bytes, g.rb, cpc_shift, sbias1, ta, i_row, cc * 16, sw, true);
          }
          cc += 2;
          while (cc >= cps) { cc -= cps; ++s; }
        }
      }
      fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(t_full, 0u, (uint32_t)crank); else mbar_arrive(t_full); }   // CTA pair: the leader's MMA thread waits
      if (warp == 2) L2S_TRACE(2, it_no, 1);
      publish_item();                    // the previous item's stores went out a whole phase 1 ago
      // ---- phase 2: D2 -> global (the wait on d2_full happens inside, after the first residual loads are issued)
      {
        const uint32_t t2 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc_cols;
        const int row_lim = dummy ? 0 : min(p.lin, q0 + g.r_out);   // rows >= r_out of a tile are not computable here
        if constexpr (EPI_TMA) {
          epilogue_item_tma<MODE>(p, g, tw, &tmRes, &tmRaw, &tmRawT, &tmAct, &tmActT,
                                  reinterpret_cast<uint8_t*>(epi_tiles) + (size_t)(warp - 2) * kTmaStageBytes,
                                  res_full + 2 * (warp - 2), ph_res, t2, b, quad, lane, d2_full, pd2,
                                  (P.trace && blockIdx.x == 0 && warp == 2 && it_no < 32) ? P.trace + 2304 + it_no * 8 : nullptr);
        } else if constexpr (DUAL) {   // 80-register budget: 16-column chunks, no residual double buffer
          epilogue_item_rows<16, MODE, false>(p, tile, t2, b, q0, row_lim, g.msub, g.c, quad, half, lane, 0, d2_full, pd2);
        } else {
          if (g.cw == 32) epilogue_item_rows<32, MODE, true>(p, tile, t2, b, q0, row_lim, g.msub, g.c, quad, half, lane, 0, d2_full, pd2);
          else epilogue_item_rows<16, MODE, true>(p, tile, t2, b, q0, row_lim, g.msub, g.c, quad, half, lane, 0, d2_full, pd2);
        }
        pd2 ^= 1u;
      }
      if (warp == 2) L2S_TRACE(2, it_no, 2);
      if (P.done_flags && !dummy) unpublished = item;   // this warp's rows of the item are issued: counted later (publish_item)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(d2_empty, 0u, (uint32_t)crank); else mbar_arrive(d2_empty); }
      if (warp == 2) L2S_TRACE(2, it_no, 3);
    }
    publish_item();
    if (EPI_TMA && lane == 0) bulk_wait_all();   // every TMA store has landed before the CTA exits
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.span && threadIdx.x == 0) atomicMax(&P.span[1], (unsigned long long)gtime());
  if (P.trace && threadIdx.x == 0 && blockIdx.x < 512) P.trace[768 + blockIdx.x * 3 + 2] = gtime();
  if (g.cluster > 1) cluster_sync_all();      // no CTA leaves while its partner may still multicast into it
  if (warp == 1) { if constexpr (CG2) tmem_dealloc_cg2(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc_dyn(tmem_base, (uint32_t)g.tmem_cols); }
}

// ------------------------------------------------------------------ host side

inline bool pair_plan_with(int c, int k, int dil, int lin, int batch, int smem_budget, bool dual, bool cluster_ok,
                           bool alias_ok, int cw_pref, int min_sb, bool epi_tma, bool force_alias, PairGeom* out) {
  PairGeom g{};
  if (c % 16 != 0 || c > 256 || k < 1 || k > kMaxTaps || (k & 1) == 0) return false;
  g.c = c; g.k = k; g.dil = dil;
  g.h1 = dil * (k - 1) / 2;
  g.h2 = (k - 1) / 2;
  g.rb = (c >= 64 ? 64 : c) * 2;
  if (g.rb != 32 && g.rb != 64 && g.rb != 128) return false;
  g.kc = c * 2 / g.rb;
  g.k16 = g.rb / 32;
  g.cw = (!dual && c % 32 == 0 && cw_pref == 32) ? 32 : 16;
  g.dual = dual ? 1 : 0;
  g.tile_words = 32 * g.cw;
  g.epi_tma = (epi_tma && dual && c <= 64) ? 1 : 0;
  if (g.epi_tma) g.tile_words = kTmaStageBytes / 4;
  int tb = 1;
  while (tb < k && tb < 16 && (tb * 2) * c * g.rb <= 16384) tb *= 2;
  if (tb > k) tb = k;
  g.tb = tb;
  g.n_tstages = (k + tb - 1) / tb;
  const int full_stage = tb * c * g.rb;
  // CTA pairs issuing cta_group::2 MMAs keep only their half of every weight stage
  const bool want_cg2 = (!dual || g_pair_cg2) && cluster_ok && g_pair_cg2 && (c == 128 || c == 256) && tb == 1 && !g.epi_tma;
  const int bar_bytes = 1024 + 512 + 1024 + 1024 + kTcEpiWarps * g.tile_words * 4;   // alignment slack, barriers, staging alignment, bias1, tiles
  int msub = (dual ? 128 : 256) / c;
  if (msub < 1) msub = 1;
  if (msub > 8) msub = 8;
  const int need = (lin + 2 * g.h2 + 127) / 128;
  if (msub > need) msub = need;
  for (; msub >= 1; --msub) {
    g.msub = msub;
    g.mt = msub * 128;
    g.r_out = g.mt - 2 * g.h2;
    if (g.r_out < 1) continue;
    const int slab_rows = g.mt + 2 * g.h1;
    g.n_loads = (slab_rows + 255) / 256;
    g.box_rows = (((slab_rows + g.n_loads - 1) / g.n_loads) + 7) & ~7;
    g.a_rows = g.n_loads * g.box_rows;
    g.slab_bytes = g.a_rows * g.rb;
    g.t_rows = (g.mt + 2 * g.h2 + 7) & ~7;
    g.t_chunk_bytes = g.t_rows * g.rb;
    // C >= 128 is bound by how many weight bytes are in flight (ring depth x stage size / L2 latency):
    // give the ring the room by letting the A-slab ring share the T-slab region.
    g.alias_at = (alias_ok && (c >= 128 || force_alias)) ? 1 : 0;
    const int t_bytes = g.kc * g.t_chunk_bytes;
    g.m_items = (lin + g.r_out - 1) / g.r_out;
    g.total_items = batch * g.m_items;
    const bool cg2_here = want_cg2 && g.total_items >= 2;
    g.bstage_bytes = cg2_here ? full_stage / 2 : full_stage;
    int sa = g.kc + 1 < 4 ? g.kc + 1 : 4, sb = 4;
    auto region = [&](int sa_) {
      const int a_bytes = sa_ * g.slab_bytes;
      return g.alias_at ? (a_bytes > t_bytes ? a_bytes : t_bytes) : a_bytes + t_bytes;
    };
    if (g.alias_at) while (sa > 2 && sa * g.slab_bytes > t_bytes) --sa;      // aliased slabs are free up to the T size
    while (sa > 2 && region(sa) + sb * g.bstage_bytes + bar_bytes > smem_budget) --sa;
    while (sb > 2 && region(sa) + sb * g.bstage_bytes + bar_bytes > smem_budget) --sb;
    if (region(sa) + sb * g.bstage_bytes + bar_bytes > smem_budget) continue;
    while (sb < kTcMaxStagesB && sb < 2 * g.n_tstages * g.kc && (sb + 1) * g.bstage_bytes <= 160 * 1024 &&
           region(sa) + (sb + 1) * g.bstage_bytes + bar_bytes <= smem_budget)
      ++sb;
    if (sb < min_sb) continue;
    g.sa = sa;
    g.sb = sb;
    g.region_bytes = (region(sa) + 1023) & ~1023;
    g.smem_bytes = g.region_bytes + sb * g.bstage_bytes + bar_bytes;
    int cols = 32;
    while (cols < 2 * msub * c) cols <<= 1;
    if (cols > (dual ? 256 : 512)) continue;
    g.tmem_cols = cols;
    g.idesc = umma_idesc_bf16(128u, (uint32_t)c);
    // weight multicast pays where weights dominate the L2 traffic and a stage is one tap (tb == 1): C >= 128
    // plain weight multicast only for one-CTA-per-SM plans; cta_group::2 pairs also for the two-CTAs-per-SM plans
    g.cluster = ((!dual || g_pair_cg2) && cluster_ok && c >= 128 && g.tb == 1 && (c / 2) % 8 == 0 && g.total_items >= 2) ? 2 : 1;
    g.cg2 = (g.cluster == 2 && cg2_here) ? 1 : 0;
    if (cg2_here && !g.cg2) continue;   // (cannot happen: the same conditions) stage size and mode must agree
    *out = g;
    return true;
  }
  return false;
}

// Two co-resident CTAs per SM when the step fits twice (C <= 64 in the shipped config): one CTA's
// MMA / phase 1 overlaps the other's load/store-heavy phase 2.  (Measured: with half the CTAs every
// stage takes ~1.7x longer, i.e. the kernels are per-SM latency bound, not chip-memory bound.)
inline bool pair_plan(int c, int k, int dil, int lin, int batch, int smem_budget, bool allow_dual, bool allow_cluster,
                      bool allow_alias, bool want_tma, bool want_tma_alias, PairGeom* out) {
  // TMA-streamed phase 2: only where its staging fits WITHOUT aliasing the A ring into the T slab (aliasing
  // exposes the ~1.5 us A-slab load latency in every item, which cancels the faster phase 2; measured)
  if (allow_dual && want_tma &&
      pair_plan_with(c, k, dil, lin, batch, 112 * 1024, true, false, allow_alias, 16, 2, true, want_tma_alias, out) && out->epi_tma)
    return true;
  if (allow_dual && pair_plan_with(c, k, dil, lin, batch, 112 * 1024, true, allow_cluster && g_pair_cg2 != 0, allow_alias, 16, 2, false, false, out)) return true;
  // single CTA per SM: 32-column epilogue chunks (full 128-byte lines) as long as >= 3 weight stages still fit,
  // else 16-column chunks (smaller transpose tiles) so the room goes to the weight ring
  if (g_pair_pref == 1 && c >= 256 &&
      pair_plan_with(c, k, dil, lin, batch, smem_budget, false, allow_cluster, allow_alias, 16, 4, false, false, out))
    return true;   // experiment: 16-column tiles, >= 4 weight stages
  if (pair_plan_with(c, k, dil, lin, batch, smem_budget, false, allow_cluster, allow_alias, 32, c >= 128 ? 3 : 2, false, false, out)) return true;
  return pair_plan_with(c, k, dil, lin, batch, smem_budget, false, allow_cluster, allow_alias, 16, 2, false, false, out);
}

struct PairEpiMaps {
  CUtensorMap res, raw, raw_tail, act, act_tail;   // only read by the EPI_TMA kernels
};

template <int MODE, bool DUAL, bool EPI_TMA, bool CG2 = false>
inline cudaError_t launch_pair_mode(const PairParams& P, const CUtensorMap& tmA, const CUtensorMap& tmW1,
                                    const CUtensorMap& tmW2, const PairEpiMaps& em, int grid, cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(pair_tc_kernel<MODE, DUAL, EPI_TMA, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pair_tc_kernel<MODE, DUAL, EPI_TMA, CG2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = (size_t)P.g.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (P.g.cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)P.g.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_tc_pdl || P.wait_flags) {   // programmatic dependent launch: this grid may start while the previous one drains
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, pair_tc_kernel<MODE, DUAL, EPI_TMA, CG2>, tmA, tmW1, tmW2, em.res, em.raw, em.raw_tail, em.act,
                            em.act_tail, P);
}

// c: the c2 epilogue description (bias = b2, res, acc_in, outputs, div, slope, lin = mrows = L, ntot = C, out_valid = L * C).
// Defined in tu_pair_tc.cu (the only translation unit that instantiates pair_tc_kernel); declared everywhere else.
#ifndef L2S_TU_PAIR_TC
cudaError_t launch_pair_tc(const ConvParams& c, const float* bias1, const PairGeom& g, const CUtensorMap& tmA,
                           const CUtensorMap& tmW1, const CUtensorMap& tmW2, const PairEpiMaps& em, int num_ctas,
                           cudaStream_t stream, long long* trace = nullptr, unsigned long long* span = nullptr,
                           const int* wait_flags = nullptr, int* done_flags = nullptr);
#else
cudaError_t launch_pair_tc(const ConvParams& c, const float* bias1, const PairGeom& g, const CUtensorMap& tmA,
                           const CUtensorMap& tmW1, const CUtensorMap& tmW2, const PairEpiMaps& em, int num_ctas,
                           cudaStream_t stream, long long* trace, unsigned long long* span, const int* wait_flags, int* done_flags) {
  PairParams P;
  P.c = c;
  P.bias1 = bias1;
  P.g = g;
  P.trace = trace;
  P.span = span;
  P.wait_flags = wait_flags;
  P.done_flags = done_flags;
  const int cap = num_ctas * (g.dual ? 2 : 1);
  int grid = g.total_items < cap ? g.total_items : cap;
  if (grid < 1) grid = 1;
  if (g.cluster > 1) {                       // CTA pairs: even grid, one pair per two items at most
    const int pairs_needed = (g.total_items + 1) / 2;
    int pairs = cap / 2 < pairs_needed ? cap / 2 : pairs_needed;
    if (pairs < 1) pairs = 1;
    grid = 2 * pairs;
  }
  const int mode = (c.res ? kEpiRes : 0) | ((c.acc_in || c.div != 1.0f) ? kEpiAcc : 0) | (c.out_raw ? kEpiRaw : 0) |
                   (c.out_act ? kEpiAct : 0);
  if (g.epi_tma) {   // asynchronous phase 2: dual plans, residual + raw (+ act)
    if (mode == 5) return launch_pair_mode<5, true, true>(P, tmA, tmW1, tmW2, em, grid, stream);
    if (mode == 13) return launch_pair_mode<13, true, true>(P, tmA, tmW1, tmW2, em, grid, stream);
    return cudaErrorInvalidValue;
  }
  if (g.cg2) {       // CTA pairs issuing cta_group::2 MMAs: its own instantiations (see pair_tc_kernel)
    switch (mode) {
#define L2S_PMODE2(m)                                                                                  \
  case m:                                                                                              \
    return g.dual ? launch_pair_mode<m, true, false, true>(P, tmA, tmW1, tmW2, em, grid, stream)       \
                  : launch_pair_mode<m, false, false, true>(P, tmA, tmW1, tmW2, em, grid, stream);
      L2S_PMODE2(5) L2S_PMODE2(7) L2S_PMODE2(11) L2S_PMODE2(13) L2S_PMODE2(15)
#undef L2S_PMODE2
      default: return cudaErrorInvalidValue;
    }
  }
  switch (mode) {
#define L2S_PMODE(m)                                                                          \
  case m:                                                                                     \
    return g.dual ? launch_pair_mode<m, true, false>(P, tmA, tmW1, tmW2, em, grid, stream)    \
                  : launch_pair_mode<m, false, false>(P, tmA, tmW1, tmW2, em, grid, stream);
    L2S_PMODE(5) L2S_PMODE(7) L2S_PMODE(11) L2S_PMODE(13) L2S_PMODE(15)
#undef L2S_PMODE
    default: return cudaErrorInvalidValue;   // a ResBlock step always has the residual and an output
  }
}
#endif  // L2S_TU_PAIR_TC

}  // namespace l2s
