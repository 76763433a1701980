// tcgen05 implementation of the tap-offset convolution (conv_common.cuh), bf16
// operands, fp32 accumulation in TMEM.
//
// GEMM view per work item:  D[128*msub rows (time), nt cols (out channels)]
//     += sum over taps j, input channels ci of  A_j[row, ci] * W_j[col, ci]
// where A_j is the SAME channels-last activation tile shifted by tap_off[j] rows.
//
//   * A: ONE halo slab per (item, 64-channel K chunk) is brought in by TMA
//     (rows [q0+min_off, q0+min_off+slab_rows), zero filled outside [0, lin) by the
//     TMA unit, which is exactly the per-layer zero padding of the reference).
//     Every tap is then a row-shifted shared-memory matrix descriptor over that
//     slab: k taps cost one global->shared transfer, not k.
//   * W: [ntaps][ntot][cin] bf16, K-major; a pipeline stage holds `tb` taps of one
//     K chunk for one nt-wide column tile.
//   * D: msub accumulators of 128 x nt fp32 in TMEM, double buffered so the
//     epilogue of item i overlaps the MMAs of item i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer (one elected lane), warps 2..9 = epilogue (TMEM -> registers -> fused
// bias / residual / branch-sum / mean / leaky-ReLU -> global).
// Persistent: grid = min(items, SMs), static round-robin over items.
#pragma once
#include <cuda.h>

#include "conv_common.cuh"
#include "ptx.cuh"

namespace l2s {

constexpr int kTcThreads = 320;
// host-side switch: launch the tcgen05 kernels with programmatic stream serialization (knob `pdl`)
inline int g_tc_pdl = 0;   // measured neutral on B200 (secondary CTAs cannot start before primaries free their shared memory)
constexpr int kTcEpiWarps = 8;
constexpr int kTcMaxStagesA = 8;
constexpr int kTcMaxStagesB = 8;

struct TcGeom {
  int esz;           // operand element size: 2 (bf16) or 4 (tf32 mode: fp32 words)
  int rb;            // bytes per shared-memory operand row: 32 / 64 / 128 (= swizzle span)
  int kc;            // K chunks per tap (cin_pad * 2 / rb)
  int k16;           // MMAs (K = 16) per chunk (rb / 32)
  int nt;            // column tile (MMA N), multiple of 16, <= 256
  int n_ntiles;      // ntot / nt
  int msub;          // 128-row accumulators per item
  int m_items;       // ceil(mrows / (128 * msub))
  int min_off;       // min(tap_off)
  int box_rows;      // rows per TMA box of the slab (multiple of 8, <= 256)
  int n_loads;       // TMA boxes per slab
  int slab_bytes;    // n_loads * box_rows * rb
  int sa;            // slab ring depth
  int tb;            // taps per W stage
  int n_tstages;     // ceil(ntaps / tb)
  int bstage_bytes;  // tb * nt * rb
  int sb;            // W ring depth
  int tmem_cols;     // power of two >= 2 * msub * nt
  int cw;            // epilogue chunk width in columns: 32 when nt % 32 == 0, else 16
  int ctas_per_sm;   // 2: planned so that two CTAs fit one SM
  int total_items;
  int cg2;               // 1: CTA pairs (cluster of 2) issue cta_group::2 MMAs over two row tiles; a CTA keeps half of a W stage
  int per_tap;           // 1: probe/fallback mode, one A tile per tap by TMA (no row-shifted descriptors)
  uint32_t idesc;
  int smem_bytes;
};

struct TcParams {
  ConvParams c;
  TcGeom g;
  unsigned long long* span;   // debug: [0] min CTA start, [1] max CTA end (globaltimer), null in production
  long long* trace;   // debug: [3 roles][64 items][4] globaltimer stamps of CTA 0, then [512 CTAs][smid, t0, t1] (null in production)
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define L2S_TRACE(role, slot, ev)                                                            \
  do {                                                                                     \
    if (P.trace && blockIdx.x == 0 && (slot) < 64 && lane == 0) P.trace[((role) * 64 + (slot)) * 4 + (ev)] = gtime(); \
  } while (0)

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

constexpr int kEpiTileWords = 32 * 32;   // warp-private transpose tile: 32 rows x CW fp32, XOR-swizzled float4 slots

// Epilogue of one [32 rows x CW columns] accumulator chunk owned by one warp.
// TMEM hands every lane one ROW (32x32b shape); storing that way would make each
// 16-byte global access of the warp touch 32 different 128-byte lines.  The chunk
// is therefore transposed through a warp-private shared-memory tile so that the
// residual / branch-sum loads and the fp32 / bf16 stores are line-contiguous:
// lane -> (row = i * RPI + lane / LPR, 4 consecutive columns (lane % LPR) * 4).
// Tile slot of (row, float4 c4): row * LPR + (c4 ^ swz(row)); both the row-per-lane
// writes and the column-per-lane reads are bank-conflict free.
template <int CW>
__device__ __forceinline__ int epi_slot(int row, int c4) {
  if constexpr (CW == 32) return row * 8 + (c4 ^ (row & 7));
  else return row * 4 + (c4 ^ ((row >> 1) & 3));
}

// Where one chunk of the current item lands in global memory, as seen by this lane.
struct EpiChunk {
  long long e0;      // element index (whole batch) of my granule in the chunk's first row group
  uint32_t taddr;    // TMEM address of the chunk (my quadrant's lanes, first column)
  uint32_t okmask;   // bit i: row group i is inside the utterance / valid output range
  int n;             // first of my 4 output columns
  float4 bv;         // bias of my 4 columns, fetched at locate time (its latency hides behind the TMEM load)
};

template <int CW>
__device__ __forceinline__ EpiChunk epi_locate(const ConvParams& p, uint32_t taddr, int b, int q_base, int n_base,
                                               int crow, int c4, int mrows_eff, int row_lo = 0) {
  constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
  EpiChunk c;
  c.n = n_base + c4 * 4;
  c.taddr = taddr;
  c.bv = __ldg(reinterpret_cast<const float4*>(p.bias + c.n));
  // Element index (within the utterance) of my granule in row q_base + crow; rows advance by
  // RPI * ntot.  Valid row groups are those with q < mrows and 0 <= idx < out_valid; idx grows
  // with i, so they form one contiguous range [i_lo, i_hi).
  const int q0 = q_base + crow;
  if (p.out_shift == 0 && p.mrows == p.lin) {
    // plain convolutions (everything but the polyphase ConvTranspose1d): validity is a row test
    c.e0 = ((long long)b * p.lin + q0) * p.ntot + c.n;      // out_valid == lin * ntot
    const int left = mrows_eff - q0;                          // valid rows from q0 on
    const int n_ok = left <= 0 ? 0 : (left + RPI - 1) / RPI;
    c.okmask = n_ok >= ITERS ? (1u << ITERS) - 1u : (1u << n_ok) - 1u;
    // rows below row_lo (the halo rows of a fused ResBlock tile) are not stored either
    const int below = row_lo - q0;
    if (below > 0) {
      const int n_lo = (below + RPI - 1) / RPI;
      c.okmask &= n_lo >= ITERS ? 0u : ~((1u << n_lo) - 1u);
    }
    return c;
  }
  const long long idx0 = (long long)q0 * p.ntot + c.n + p.out_shift;
  const int step = RPI * p.ntot;
  c.e0 = (long long)b * p.out_valid + idx0;
  if (idx0 >= 0 && idx0 + (long long)(ITERS - 1) * step < p.out_valid && q0 + (ITERS - 1) * RPI < mrows_eff) {
    c.okmask = (1u << ITERS) - 1u;
  } else {
    int i_lo = 0, i_hi = mrows_eff > q0 ? (mrows_eff - q0 + RPI - 1) / RPI : 0;
    if (idx0 < 0) i_lo = (int)((-idx0 + step - 1) / step);
    const long long room = p.out_valid - idx0;
    const int lim = room <= 0 ? 0 : (int)((room + step - 1) / step);
    i_hi = i_hi < lim ? i_hi : lim;
    i_hi = i_hi < ITERS ? i_hi : ITERS;
    i_lo = i_lo < ITERS ? i_lo : ITERS;
    c.okmask = i_hi > i_lo ? (((1u << i_hi) - 1u) & ~((1u << i_lo) - 1u)) : 0u;
  }
  return c;
}

// Epilogue mode bits (compile-time: dead paths disappear from the instruction stream).
constexpr int kEpiRes = 1, kEpiAcc = 2, kEpiRaw = 4, kEpiAct = 8;

// Residual-stream / branch-sum loads of a chunk: plain global loads that do not depend on the
// accumulator, so they are issued as early as possible (one chunk ahead, and before waiting for the MMAs).
template <int CW, int MODE>
__device__ __forceinline__ void epi_load_res(const ConvParams& p, const EpiChunk& c, float4 (&rv)[CW / 4]) {
  constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
  if constexpr ((MODE & kEpiRes) != 0) {
    const float* rp = p.res + c.e0;
    const int step = RPI * p.ntot;
    if (p.pf & 4) {     // chained steps: the rows were written by a grid that may still be running -- read them at L2, never from L1
#pragma unroll
      for (int i = 0; i < ITERS; ++i, rp += step)
        rv[i] = ((c.okmask >> i) & 1u) ? __ldcg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
      return;
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i, rp += step)
      rv[i] = ((c.okmask >> i) & 1u) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int CW, int MODE>
__device__ __forceinline__ void epi_load_acc(const ConvParams& p, const EpiChunk& c, float4 (&av)[CW / 4]) {
  constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
  if constexpr ((MODE & kEpiAcc) != 0) {
    if (p.acc_in) {
      const float* ap = p.acc_in + c.e0;
      const int step = RPI * p.ntot;
#pragma unroll
      for (int i = 0; i < ITERS; ++i, ap += step)
        av[i] = ((c.okmask >> i) & 1u) ? *reinterpret_cast<const float4*>(ap) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < ITERS; ++i) av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// The same lines pulled into L1 ahead of time, holding no registers (80-register kernels cannot double-buffer the
// loads above; without this every chunk pays a full memory latency).  One lane per row segment issues.
template <int CW, int MODE>
__device__ __forceinline__ void epi_prefetch(const ConvParams& p, const EpiChunk& c, int c4) {
  constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
  if (c4 != 0) return;
  const int step = RPI * p.ntot;
  if constexpr ((MODE & kEpiRes) != 0) {
    if (!(p.pf & 4)) {
#pragma unroll
      for (int i = 0; i < ITERS; ++i)
        if ((c.okmask >> i) & 1u) prefetch_l1(p.res + c.e0 + (long long)i * step);
    }
  }
  if constexpr ((MODE & kEpiAcc) != 0) {
    if (p.acc_in) {
#pragma unroll
      for (int i = 0; i < ITERS; ++i)
        if ((c.okmask >> i) & 1u) prefetch_l1(p.acc_in + c.e0 + (long long)i * step);
    }
  }
}

// Stage 2: accumulator rows -> transpose tile (one row per lane, swizzled float4 slots).
template <int CW>
__device__ __forceinline__ void epi_stage(float4* tile4, uint32_t taddr, int lane) {
  uint32_t r[CW];
  if constexpr (CW == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
  tmem_ld_wait();      // kept adjacent to the load: see epilogue_item_rows
#pragma unroll
  for (int j = 0; j < CW / 4; ++j)
    tile4[epi_slot<CW>(lane, j)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
}

// Stage 3: read back column-per-lane, apply bias / residual / branch sum / mean / leaky-ReLU, store.
template <int CW, int MODE, bool FULL>
__device__ __forceinline__ void epi_finish(const ConvParams& p, const EpiChunk& c, const float4* tile4,
                                           const float4 (&rv)[CW / 4], const float4 (&av)[CW / 4], int crow, int c4) {
  constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
  const int step = RPI * p.ntot;
  const float4 bv = c.bv;
  float* raw_p = p.out_raw + c.e0;
  __nv_bfloat16* act_p = reinterpret_cast<__nv_bfloat16*>(p.out_act) + c.e0;
  const __nv_bfloat162 slope2 = __float2bfloat162_rn(p.slope);
  // The tensor-core (bf16) mode multiplies by the reciprocal of the branch count; the
  // fp32 CUDA-core mode keeps the reference's true division (conv_common.cuh).
  const float inv_div = 1.0f / p.div;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const float4 t = tile4[epi_slot<CW>(i * RPI + crow, c4)];
    const bool ok = FULL || ((c.okmask >> i) & 1u);
    float v0 = t.x + bv.x, v1 = t.y + bv.y, v2 = t.z + bv.z, v3 = t.w + bv.w;
    if constexpr ((MODE & kEpiRes) != 0) { v0 += rv[i].x; v1 += rv[i].y; v2 += rv[i].z; v3 += rv[i].w; }
    if constexpr ((MODE & kEpiAcc) != 0) {
      v0 += av[i].x; v1 += av[i].y; v2 += av[i].z; v3 += av[i].w;
      v0 *= inv_div; v1 *= inv_div; v2 *= inv_div; v3 *= inv_div;   // div != 1 only on the branch-mean conv
    }
    if (ok) {
      if constexpr ((MODE & kEpiRaw) != 0)
        *reinterpret_cast<float4*>(raw_p + (long long)i * step) = make_float4(v0, v1, v2, v3);
      if constexpr ((MODE & kEpiAct) != 0) {
        uint2 pk;   // leaky_relu for 0 < slope < 1 is max(v, v * slope)
        if (p.act_f32) {   // tf32 mode: the activated copy stays fp32
          // rounded to nearest tf32 here: the tensor core would otherwise TRUNCATE the low 13 mantissa bits
          const float sl = p.slope;
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out_act) + c.e0 + (long long)i * step) =
              make_float4(round_tf32(fmaxf(v0, v0 * sl)), round_tf32(fmaxf(v1, v1 * sl)), round_tf32(fmaxf(v2, v2 * sl)),
                          round_tf32(fmaxf(v3, v3 * sl)));
        } else {
          pk.x = lrelu_bf16x2(v0, v1, slope2);
          pk.y = lrelu_bf16x2(v2, v3, slope2);
          *reinterpret_cast<uint2*>(act_p + (long long)i * step) = pk;
        }
      }
    }
  }
}

// All chunks of one item owned by this warp (quadrant `quad`, every second chunk starting at
// `half`): msub accumulators of 128 rows x nt columns at TMEM address t_base; tile row 0 is output
// row q0 of utterance b, column 0 is output column n_tile_base; rows >= row_lim are not stored.
// `bar` is the accumulator-ready barrier: the first chunk's residual loads are issued BEFORE
// waiting on it, and chunk j+1's residual loads while chunk j is finished (two chunks of
// residual in flight per warp; with one, exposed DRAM latency capped a CTA at ~3 TB/s / 148).
// The TMEM load and its wait stay adjacent: a tcgen05.ld left in flight across other code is not
// safe (the compiler may move its destination registers before wait::ld; measured wrong results).
template <int CW, int MODE, bool PIPE>
__device__ __forceinline__ void epilogue_item_rows(const ConvParams& p, float* tile, uint32_t t_base, int b, int q0,
                                                   int row_lim, int msub, int nt, int quad, int half, int lane,
                                                   int n_tile_base, uint64_t* bar, uint32_t parity, int row_lo = 0) {
  constexpr int LPR = CW / 4;
  constexpr uint32_t kAll = (1u << (CW / 4)) - 1u;
  // PIPE: double-buffer the residual registers (needs the 168-register budget of the fused kernel)
  constexpr bool kPipe = PIPE && (MODE & kEpiRes) != 0 && (MODE & kEpiAcc) == 0;
  const int crow = lane / LPR;
  const int c4 = lane % LPR;
  float4* tile4 = reinterpret_cast<float4*>(tile);
  const int cps = nt / CW;                    // chunks per 128-row accumulator
  auto locate = [&](int s_, int cc_) {
    return epi_locate<CW>(p, t_base + (uint32_t)(s_ * nt + cc_ * CW), b, q0 + s_ * 128 + quad * 32, n_tile_base + cc_ * CW,
                          crow, c4, row_lim, row_lo);
  };
  auto finish = [&](const EpiChunk& c, const float4 (&rv)[CW / 4], const float4 (&av)[CW / 4]) {
    if (__all_sync(0xffffffffu, c.okmask == kAll)) epi_finish<CW, MODE, true>(p, c, tile4, rv, av, crow, c4);
    else epi_finish<CW, MODE, false>(p, c, tile4, rv, av, crow, c4);
  };
  int s = 0, cc = half;
  while (cc >= cps) { cc -= cps; ++s; }
  bool have = s < msub;
  EpiChunk ca{}, cb{};
  float4 rva[CW / 4], rvb[CW / 4], av[CW / 4];
  if (have) { ca = locate(s, cc); epi_load_res<CW, MODE>(p, ca, rva); }
  mbar_wait(bar, parity);
  tc_fence_after();
  while (have) {
    // ---- chunk A
    cc += 2;
    while (cc >= cps) { cc -= cps; ++s; }
    bool more = s < msub;
    epi_load_acc<CW, MODE>(p, ca, av);
    if (!kPipe && more && (p.pf & 2)) epi_prefetch<CW, MODE>(p, locate(s, cc), c4);   // experiment: next chunk's lines -> L1
    epi_stage<CW>(tile4, ca.taddr, lane);
    if (kPipe && more) { cb = locate(s, cc); epi_load_res<CW, MODE>(p, cb, rvb); }
    __syncwarp();
    finish(ca, rva, av);
    __syncwarp();   // the tile is rewritten by the next chunk
    if (!more) break;
    if (!kPipe) { ca = locate(s, cc); epi_load_res<CW, MODE>(p, ca, rva); continue; }
    // ---- chunk B (residual already in flight)
    cc += 2;
    while (cc >= cps) { cc -= cps; ++s; }
    have = s < msub;
    epi_stage<CW>(tile4, cb.taddr, lane);
    if (have) { ca = locate(s, cc); epi_load_res<CW, MODE>(p, ca, rva); }
    __syncwarp();
    finish(cb, rvb, av);
    __syncwarp();
  }
}

// K16 consecutive K = 16 slices of one (tap, 64-channel chunk): descriptors advance by 32 bytes.
template <int K16, bool CG2 = false>
__device__ __forceinline__ void issue_chunk(bool leader, uint32_t d_addr, uint32_t desc_hi, uint32_t a_lo, uint32_t b_lo,
                                            uint32_t idesc, uint32_t first, bool tf32 = false) {
#pragma unroll
  for (int k = 0; k < K16; ++k) {
    const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
    const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2u * k);
    if (leader) {
      if constexpr (CG2) umma_bf16_cg2(d_addr, da, db, idesc, (first | (uint32_t)k) != 0u ? 1u : 0u);   // bf16 plans only
      else if (tf32) umma_tf32(d_addr, da, db, idesc, (first | (uint32_t)k) != 0u ? 1u : 0u);
      else umma_bf16(d_addr, da, db, idesc, (first | (uint32_t)k) != 0u ? 1u : 0u);
    }
  }
}

#ifndef L2S_TC_MAXNREG
#define L2S_TC_MAXNREG 80
#endif
// CG2 (own instantiations: such code cannot be launched without a cluster): CTA pairs walk two neighbouring row tiles
// of the same column tile in lockstep; the leader's MMA thread issues cta_group::2 MMAs (M = 256), each CTA loads its own
// A slab and HALF of every W stage, completions are counted on the leader's barriers, commits are multicast to both.
template <int MODE, bool CG2 = false>
__global__ void __maxnreg__(L2S_TC_MAXNREG)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = P.c;
  const TcGeom& g = P.g;

  // carve dynamic shared memory (1024-byte aligned base for the swizzle pattern)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slabA = smem;
  uint8_t* stageB = slabA + (size_t)g.sa * g.slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageB + (size_t)g.sb * g.bstage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kTcMaxStagesA;
  uint64_t* b_full = a_empty + kTcMaxStagesA;
  uint64_t* b_empty = b_full + kTcMaxStagesB;
  uint64_t* acc_full = b_empty + kTcMaxStagesB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_tiles = reinterpret_cast<float*>(bars + 40);   // 36 barriers + slot, 16-byte aligned

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
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < g.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < g.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], (CG2 ? 2 : 1) * kTcEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(tmem_slot, (uint32_t)g.tmem_cols); else tmem_alloc_dyn(tmem_slot, (uint32_t)g.tmem_cols); }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();      // the partner's barriers exist before anything is signalled to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the prologue above may overlap the previous kernel's tail; its data is touched only after this wait
  pdl_launch_dependents();
  pdl_wait_prior_grid();

  const int items_per_b = g.m_items * g.n_ntiles;
  const int acc_cols = g.msub * g.nt;
  // work walk.  Unpaired: item = (b, mi, ni).  CTA pairs: pair index = (b, mp, ni), rank r takes row tile mi = 2 mp + r
  // (a row tile past the end reads zero-filled rows and stores nothing).
  const int crank = CG2 ? (int)cluster_ctarank() : 0;
  const int walkers = CG2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int walk0 = CG2 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int pairs_per_b = ((g.m_items + 1) / 2) * g.n_ntiles;
  const int walk_n = CG2 ? P.c.batch * pairs_per_b : g.total_items;
  auto decode = [&](int wk, int& b, int& mi, int& ni) {
    if constexpr (CG2) {
      b = wk / pairs_per_b;
      const int rem = wk - b * pairs_per_b;
      const int mp = rem / g.n_ntiles;
      ni = rem - mp * g.n_ntiles;
      mi = 2 * mp + crank;
    } else {
      b = wk / items_per_b;
      const int rem = wk - b * items_per_b;
      mi = rem / g.n_ntiles;
      ni = rem - mi * g.n_ntiles;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loops in uniform control flow (so addresses and
    // coordinates live in uniform registers); one elected lane issues.
    const bool leader = elect_one();
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t box_bytes = (uint32_t)(g.box_rows * g.rb);
    int it_no = 0;
    for (int wk = walk0; wk < walk_n; wk += walkers, ++it_no) {
      int b, mi, ni;
      decode(wk, b, mi, ni);
      const int row0 = mi * 128 * g.msub + g.min_off;
      for (int kc = 0; kc < g.kc; ++kc) {
        const int ch0 = kc * (g.rb / g.esz);
        if (!g.per_tap) {
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
        }
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          if (g.per_tap) {
            mbar_wait(&a_empty[ia], pa ^ 1u);
            if (leader) {
              mbar_expect_tx(&a_full[ia], (uint32_t)g.slab_bytes);
              uint8_t* dst = slabA + (size_t)ia * g.slab_bytes;
              const int r0 = row0 - g.min_off + p.tap_off[ts];
              for (int l = 0; l < g.n_loads; ++l)
                tma_load_3d(dst + (size_t)l * box_bytes, &tmA, &a_full[ia], ch0, r0 + l * g.box_rows, b);
            }
            if (++ia == g.sa) { ia = 0; pa ^= 1u; }
          }
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if (leader) {
            if constexpr (CG2) {   // this CTA's half of the stage's output-channel rows
              if (crank == 0) mbar_expect_tx(&b_full[ib], 2u * (uint32_t)g.bstage_bytes);
              tma_load_3d_cg2(stageB + (size_t)ib * g.bstage_bytes, &tmW, &b_full[ib], ch0, ni * g.nt + crank * (g.nt / 2), ts * g.tb);
            } else {
              mbar_expect_tx(&b_full[ib], (uint32_t)g.bstage_bytes);
              tma_load_3d(stageB + (size_t)ib * g.bstage_bytes, &tmW, &b_full[ib], ch0, ni * g.nt, ts * g.tb);
            }
          }
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    // Uniform control flow for the whole warp; only the tcgen05.mma / commit
    // instructions are predicated on the elected lane.  Descriptors differ only in
    // their low word (start address >> 4), advanced by additions.
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template((uint32_t)g.rb);
    const uint32_t desc_hi = (uint32_t)(tmpl >> 32);
    const uint32_t desc_lo_fixed = (uint32_t)tmpl;            // LBO field
    const uint32_t sub_step = (uint32_t)(128 * g.rb) >> 4;   // next 128-row accumulator
    const uint32_t tapw_step = (uint32_t)((CG2 ? g.nt / 2 : g.nt) * g.rb) >> 4; // next tap inside a W stage
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t pacc0 = 0, pacc1 = 0;
    int buf = 0;
    int it_no = 0;
    auto commit = [&](uint64_t* bar) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); };
    for (int wk = (CG2 && crank != 0) ? walk_n : walk0; wk < walk_n; wk += walkers, ++it_no) {   // pair: the leader issues for both
      L2S_TRACE(1, it_no, 0);
      mbar_wait(&acc_empty[buf], (buf ? pacc1 : pacc0) ^ 1u);
      L2S_TRACE(1, it_no, 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(buf * acc_cols);
      for (int kc = 0; kc < g.kc; ++kc) {
        uint32_t a_lo = 0;
        if (!g.per_tap) {
          mbar_wait(&a_full[ia], pa);
          if (kc == 0) L2S_TRACE(1, it_no, 2);
          tc_fence_after();
          a_lo = desc_lo_fixed | ((smem_u32(slabA + (size_t)ia * g.slab_bytes) & 0x3FFFFu) >> 4);
        }
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          if (g.per_tap) {
            mbar_wait(&a_full[ia], pa);
            tc_fence_after();
            a_lo = desc_lo_fixed | ((smem_u32(slabA + (size_t)ia * g.slab_bytes) & 0x3FFFFu) >> 4);
          }
          mbar_wait(&b_full[ib], pb);
          tc_fence_after();
          uint32_t b_lo = desc_lo_fixed | ((smem_u32(stageB + (size_t)ib * g.bstage_bytes) & 0x3FFFFu) >> 4);
          const int t_end = min(g.tb, p.ntaps - ts * g.tb);
          for (int t = 0; t < t_end; ++t, b_lo += tapw_step) {
            const int tap = ts * g.tb + t;
            uint32_t a_sub = a_lo + (g.per_tap ? 0u : ((uint32_t)((p.tap_off[tap] - g.min_off) * g.rb) >> 4));
            const uint32_t first = (uint32_t)(kc | tap);
            uint32_t d_addr = d_base;
            if (g.k16 == 4) {
              for (int sub = 0; sub < g.msub; ++sub, a_sub += sub_step, d_addr += (uint32_t)g.nt)
                issue_chunk<4, CG2>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first, g.esz == 4);
            } else if (g.k16 == 2) {
              for (int sub = 0; sub < g.msub; ++sub, a_sub += sub_step, d_addr += (uint32_t)g.nt)
                issue_chunk<2, CG2>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first, g.esz == 4);
            } else {
              for (int sub = 0; sub < g.msub; ++sub, a_sub += sub_step, d_addr += (uint32_t)g.nt)
                issue_chunk<1, CG2>(leader, d_addr, desc_hi, a_sub, b_lo, g.idesc, first, g.esz == 4);
            }
          }
          if (leader) commit(&b_empty[ib]);   // W stage free once these MMAs retire
          if (++ib == g.sb) { ib = 0; pb ^= 1u; }
          if (g.per_tap) {
            if (leader) commit(&a_empty[ia]);
            if (++ia == g.sa) { ia = 0; pa ^= 1u; }
          }
        }
        if (!g.per_tap) {
          if (leader) commit(&a_empty[ia]);    // slab free
          if (++ia == g.sa) { ia = 0; pa ^= 1u; }
        }
      }
      if (leader) commit(&acc_full[buf]);      // accumulators complete -> epilogue
      L2S_TRACE(1, it_no, 3);
      if (buf) pacc1 ^= 1u; else pacc0 ^= 1u;
      buf ^= 1;
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;     // two warps share a quadrant, alternate column chunks
    float* tile = epi_tiles + (size_t)(warp - 2) * kEpiTileWords;
    uint32_t pacc0 = 0, pacc1 = 0;
    int buf = 0;
    int it_no = 0;
    for (int wk = walk0; wk < walk_n; wk += walkers, ++it_no) {
      int b, mi, ni;
      decode(wk, b, mi, ni);
      if (warp == 2) L2S_TRACE(2, it_no, 0);
      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols);
      const uint32_t par = buf ? pacc1 : pacc0;
      if (g.cw == 32)
        epilogue_item_rows<32, MODE, false>(p, tile, t_base, b, mi * g.msub * 128, p.mrows, g.msub, g.nt, quad, half, lane, ni * g.nt,
                                     &acc_full[buf], par);
      else
        epilogue_item_rows<16, MODE, false>(p, tile, t_base, b, mi * g.msub * 128, p.mrows, g.msub, g.nt, quad, half, lane, ni * g.nt,
                                     &acc_full[buf], par);
      if (warp == 2) L2S_TRACE(2, it_no, 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(&acc_empty[buf], 0u, (uint32_t)crank); else mbar_arrive(&acc_empty[buf]); }
      if (warp == 2) L2S_TRACE(2, it_no, 2);
      if (buf) pacc1 ^= 1u; else pacc0 ^= 1u;
      buf ^= 1;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (P.span && threadIdx.x == 0) atomicMax(&P.span[1], (unsigned long long)gtime());
  if (P.trace && threadIdx.x == 0 && blockIdx.x < 512) P.trace[768 + blockIdx.x * 3 + 2] = gtime();
  if constexpr (CG2) cluster_sync_all();      // no CTA leaves while its partner may still signal it
  if (warp == 1) { if constexpr (CG2) tmem_dealloc_cg2(tmem_base, (uint32_t)g.tmem_cols); else tmem_dealloc_dyn(tmem_base, (uint32_t)g.tmem_cols); }
}

// ------------------------------------------------------------------ host side

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

// tensor [d2][d1][d0] (d0 contiguous) of bf16 (esz 2) or fp32 (esz 4), box [b2][b1][b0], swizzle span = b0 * esz bytes.
inline bool make_tmap_3d(CUtensorMap* out, const void* base, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                         uint32_t b1, uint32_t b2) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {d0 * (uint64_t)esz, d0 * d1 * (uint64_t)esz};
  const cuuint32_t box[3] = {b0, b1, b2};
  const cuuint32_t estr[3] = {1, 1, 1};
  const uint32_t rb = b0 * (uint32_t)esz;
  const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                           : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(out, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
inline bool make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                              uint32_t b1, uint32_t b2) {
  return make_tmap_3d(out, base, 2, d0, d1, d2, b0, b1, b2);
}

struct TcTune {
  int max_msub = 8;
  int slab_cap = 40960;        // bytes per slab
  int smem_budget = 220 * 1024;
  int per_tap = 0;
  int sa_min = 0;              // force at least this many slab ring slots when they fit
  int dual = 1;                // try the two-CTAs-per-SM plan first
  int max_ctas = 0;            // 0: number of SMs
  int max_nt = 256;            // widest column tile (MMA N)
  int cg2 = 1;                 // one-CTA-per-SM bf16 plans run as CTA pairs issuing cta_group::2 MMAs (knob tc_cg2)
};

// Shape-only planning (no device pointers): valid for any batch with the same (lin, mrows).
// acc_cols_cap bounds msub * nt (one accumulator buffer), smem_budget the dynamic shared memory.
inline bool tc_plan_with(const ConvParams& c, int batch, const TcTune& tune, int acc_cols_cap, int smem_budget,
                         TcGeom* out) {
  TcGeom g{};
  if (c.cin_pad % 16 != 0 || c.ntot % 16 != 0 || c.ntaps < 1 || c.ntaps > kMaxTaps) return false;
  g.esz = c.act_f32 ? 4 : 2;
  const int row_elems = 128 / g.esz;                      // channels per 128-byte operand row
  g.rb = (c.cin_pad >= row_elems ? row_elems : c.cin_pad) * g.esz;
  if (g.rb != 32 && g.rb != 64 && g.rb != 128) return false;
  if ((c.cin_pad * g.esz) % g.rb != 0) return false;
  g.kc = c.cin_pad * g.esz / g.rb;
  g.k16 = g.rb / 32;
  g.nt = c.ntot <= 256 ? c.ntot : 256;
  if (g.nt > tune.max_nt && tune.max_nt >= 16) g.nt = tune.max_nt;
  while (c.ntot % g.nt != 0) g.nt -= 16;
  if (g.nt > acc_cols_cap) return false;
  g.n_ntiles = c.ntot / g.nt;
  int mn = c.tap_off[0], mx = c.tap_off[0];
  for (int j = 1; j < c.ntaps; ++j) { mn = c.tap_off[j] < mn ? c.tap_off[j] : mn; mx = c.tap_off[j] > mx ? c.tap_off[j] : mx; }
  g.min_off = mn;
  const int span = tune.per_tap ? 0 : mx - mn;
  g.per_tap = tune.per_tap;
  int msub = acc_cols_cap / g.nt;
  if (msub < 1) msub = 1;
  if (msub > tune.max_msub) msub = tune.max_msub;
  const int need = (c.mrows + 127) / 128;
  if (msub > need) msub = need;
  while (msub > 1 && (msub * 128 + span) * g.rb > tune.slab_cap) --msub;
  g.msub = msub;
  g.m_items = (c.mrows + 128 * msub - 1) / (128 * msub);
  const int slab_rows = msub * 128 + span;
  g.n_loads = (slab_rows + 255) / 256;
  g.box_rows = (((slab_rows + g.n_loads - 1) / g.n_loads) + 7) & ~7;
  g.slab_bytes = g.n_loads * g.box_rows * g.rb;
  // W stage: a few taps per stage when a single tap is tiny
  int tb = 1;
  while (!tune.per_tap && tb < c.ntaps && tb < 16 && (tb * 2) * g.nt * g.rb <= 16384) tb *= 2;
  if (tb > c.ntaps) tb = c.ntaps;
  g.tb = tb;
  g.n_tstages = (c.ntaps + tb - 1) / tb;
  // CTA pairs (cta_group::2): bf16 only, one CTA per SM (the two-CTAs-per-SM plans are the HBM-bound narrow layers),
  // each CTA keeps half of a W stage
  g.cg2 = (tune.cg2 && !tune.per_tap && g.esz == 2 && (acc_cols_cap == 256 || tune.cg2 >= 2) && g.nt % 32 == 0 && batch * g.m_items >= 2) ? 1 : 0;
  g.bstage_bytes = tb * (g.cg2 ? g.nt / 2 : g.nt) * g.rb;
  // ring depths within the shared-memory budget
  const int bar_bytes = 1024 + 320 + kTcEpiWarps * kEpiTileWords * 4;  // alignment slack, barriers + TMEM slot, epilogue tiles
  int sa = g.kc + 1 < kTcMaxStagesA ? g.kc + 1 : kTcMaxStagesA;
  if (tune.per_tap) sa = 4;
  if (tune.sa_min > sa) sa = tune.sa_min < kTcMaxStagesA ? tune.sa_min : kTcMaxStagesA;
  if (sa < 2) sa = 2;
  int sb = 4;
  while (sa > 2 && sa * g.slab_bytes + sb * g.bstage_bytes + bar_bytes > smem_budget) --sa;
  while (sb > 2 && sa * g.slab_bytes + sb * g.bstage_bytes + bar_bytes > smem_budget) --sb;
  if (sa * g.slab_bytes + sb * g.bstage_bytes + bar_bytes > smem_budget) return false;
  while (sb < kTcMaxStagesB && sb < g.n_tstages * g.kc &&
         sa * g.slab_bytes + (sb + 1) * g.bstage_bytes + bar_bytes <= smem_budget && (sb + 1) * g.bstage_bytes <= 96 * 1024)
    ++sb;
  g.sa = sa;
  g.sb = sb;
  g.smem_bytes = sa * g.slab_bytes + sb * g.bstage_bytes + bar_bytes;
  int cols = 32;
  while (cols < 2 * msub * g.nt) cols <<= 1;
  if (cols > 512) return false;
  g.tmem_cols = cols;
  g.total_items = batch * g.m_items * g.n_ntiles;
  g.idesc = g.esz == 4 ? umma_idesc_tf32(128u, (uint32_t)g.nt) : umma_idesc_bf16(g.cg2 ? 256u : 128u, (uint32_t)g.nt);
  g.cw = g.nt % 32 == 0 ? 32 : 16;
  g.ctas_per_sm = 1;
  *out = g;
  return true;
}

// Two co-resident CTAs per SM (TMEM 2 x 256 columns, 2 x <= 110 KB shared memory, <= 96
// registers) hide the epilogue's latency chains on the layers where they fit (C <= 64);
// everything else gets the whole SM.
inline bool tc_plan(const ConvParams& c, int batch, const TcTune& tune, TcGeom* out) {
  if (tune.dual && !c.act_f32) {
    TcTune t = tune;
    for (; t.max_msub >= 1; t.max_msub >>= 1) {   // shrink the item until two CTAs' shared memory fits
      if (tc_plan_with(c, batch, t, 128, 110 * 1024, out) && out->tmem_cols <= 256) {
        out->ctas_per_sm = 2;
        return true;
      }
    }
  }
  return tc_plan_with(c, batch, tune, 256, tune.smem_budget, out);
}

template <int MODE, bool CG2 = false>
inline cudaError_t launch_conv_tc_mode(const TcParams& P, const CUtensorMap& tmA, const CUtensorMap& tmW, int grid,
                                       cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    // all of the SM's unified L1/shared storage as shared memory, so that two ~110 KB CTAs co-reside
    e = cudaFuncSetAttribute(conv_tc_kernel<MODE, CG2>, cudaFuncAttributePreferredSharedMemoryCarveout,
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
  unsigned na = 0;
  if (CG2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_tc_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE, CG2>, tmA, tmW, P);
}

// Defined in tu_conv_tc.cu (the only translation unit that instantiates conv_tc_kernel); declared everywhere else.
#ifndef L2S_TU_CONV_TC
cudaError_t launch_conv_tc(const ConvParams& c, const TcGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmW,
                           int num_ctas, cudaStream_t stream, long long* trace = nullptr,
                           unsigned long long* span = nullptr);
#else
cudaError_t launch_conv_tc(const ConvParams& c, const TcGeom& g, const CUtensorMap& tmA, const CUtensorMap& tmW,
                           int num_ctas, cudaStream_t stream, long long* trace, unsigned long long* span) {
  TcParams P;
  P.c = c;
  P.g = g;
  P.trace = trace;
  P.span = span;
  const int cap = num_ctas * (g.ctas_per_sm > 1 ? g.ctas_per_sm : 1);
  int grid = g.total_items < cap ? g.total_items : cap;
  if (grid < 1) grid = 1;
  if (g.cg2) {                               // CTA pairs: even grid, one pair per (b, row-tile pair, column tile)
    const int pairs_needed = c.batch * ((g.m_items + 1) / 2) * g.n_ntiles;
    int pairs = cap / 2 < pairs_needed ? cap / 2 : pairs_needed;
    if (pairs < 1) pairs = 1;
    grid = 2 * pairs;
  }
  const int mode = (c.res ? kEpiRes : 0) | ((c.acc_in || c.div != 1.0f) ? kEpiAcc : 0) | (c.out_raw ? kEpiRaw : 0) |
                   (c.out_act ? kEpiAct : 0);
  switch (mode) {
#define L2S_MODE(m) case m: return g.cg2 ? launch_conv_tc_mode<m, true>(P, tmA, tmW, grid, stream) : launch_conv_tc_mode<m>(P, tmA, tmW, grid, stream);
    L2S_MODE(4) L2S_MODE(5) L2S_MODE(6) L2S_MODE(7) L2S_MODE(8) L2S_MODE(9) L2S_MODE(10) L2S_MODE(11) L2S_MODE(12)
    L2S_MODE(13) L2S_MODE(14) L2S_MODE(15)
#undef L2S_MODE
    default: return cudaErrorInvalidValue;   // a conv with no output
  }
}
#endif  // L2S_TU_CONV_TC

}  // namespace l2s
