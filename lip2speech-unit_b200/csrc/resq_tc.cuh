// The whole-ResBlock kernel of res_tc.cuh with a SKEWED schedule for the one-CTA-per-SM plans.
//
// In res_tc_kernel the MMAs of a conv and the epilogue phase that follows it strictly alternate (one S slab, one
// "conv done" barrier): a C = 64, k = 11 tile spends 26 us in MMAs, 8 us in the six gaps between convs and 13.5 us in
// the output of the tile and the load of the next one (tools/res_trace.py).  Co-resident CTAs hide that only where the
// halo leaves room for small tiles.  Here one CTA overlaps the phases with its own MMAs:
//
//   * two S slabs: A holds lrelu(x) (input of every c1), B holds lrelu(c1 + b1) (input of every c2); a phase writes the
//     slab the running conv does not read;
//   * the tile is cut into GRANULES (one 128-row accumulator; two for C = 16) with a barrier pair each:
//     m_done[conv type][granule]  the conv's MMAs on this granule have retired        (tcgen05.commit)
//     e_done[conv type][granule]  the phase after that conv has rewritten the granule's S rows (epilogue warps)
//   * a conv still streams every weight stage ONCE, but its first `gh` and last `gt` stages are walked granule-outer:
//     the tail group commits granule after granule, so phase A / B of granule 0 starts while the tensor core still
//     works on granules 1.., and the head group of the NEXT conv starts on granule 0 as soon as that granule (and, for
//     taps right of the centre, its successor) has been rewritten.  Every accumulator still receives its taps in
//     ascending order: results are bit-identical to res_tc_kernel.
//   * tile boundary: while the last c2 of a tile runs, the epilogue warps already write lrelu(x) of the NEXT tile into
//     slab A, so the first c1 of the next tile is queued right behind that c2 and runs while the tile is written out;
//     X (TMEM) of the next tile is loaded after the output, before the first phase A releases the next c2.
//
// Barrier phases: completion i of e_done[1][*] releases c1 number i (B phase of step i - 1, or the slab-A load of a
// tile), m_done[0][*] completion i is its commit, e_done[0][*] completion i the phase A after it, m_done[1][*]
// completion i the commit of c2 number i; everybody waits with parity i & 1.
#pragma once
#include "res_tc.cuh"

namespace l2s {

constexpr int kResqMaxGran = 8;
constexpr int kResqMaxStages = 16;   // weight ring slots (res_tc_kernel: kTcMaxStagesB = 8)

// first unit >= lo owned by this warp (units half, half + ustep, ...)
__device__ __forceinline__ int resq_first_unit(const ResLane& w, int lo) {
  int r = (w.half - lo) % w.ustep;
  if (r < 0) r += w.ustep;
  return lo + r;
}

// Phase A / B on the units [u_lo, u_hi) of one granule: TMEM (D1 or X, bias included) -> leaky-ReLU -> bf16 -> slab.
template <bool EDGE>
__device__ __forceinline__ void resq_phase(const ResGeom& g, const ResLane& w, uint8_t* slab, uint32_t t_quad, int t_row0, int lin,
                                           int u_lo, int u_hi) {
  for (int u = resq_first_unit(w, u_lo); u < u_hi; u += w.ustep) {
    uint32_t r[32];
    tmem_ld32(t_quad + (uint32_t)(32 * u), r);
    tmem_ld_wait();
    int row, c0;
    res_unit_pos(g, w, u, row, c0);
    res_store_act<EDGE>(g, w, slab, r, row, c0, t_row0, lin);
  }
}

// x (global fp32) -> lrelu -> bf16 -> slab A, this warp's units of the whole tile (res_load_x without the TMEM half).
__device__ __forceinline__ void resq_load_s(const ResParams& P, const ResLane& w, uint8_t* slab, int b, int t_row0, int lin) {
  const ResGeom& g = P.g;
  const int n_units = (g.msub * g.c) >> 5;
  const bool c16 = g.c == 16;
  const float* xb = P.x + (long long)b * lin * g.c;
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
  }
}

// x (global fp32, L2 / L1 hot: resq_load_s read it a moment ago) + b2 of the first step -> X (TMEM), this warp's units.
__device__ __forceinline__ void resq_load_x(const ResParams& P, const ResLane& w, uint32_t x_quad, int b, int t_row0, const float* sbias2,
                                            int lin) {
  const ResGeom& g = P.g;
  const int n_units = (g.msub * g.c) >> 5;
  const int cmask = g.c - 1;
  const bool c16 = g.c == 16;
  const float* xb = P.x + (long long)b * lin * g.c;
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
      const float4 bq = *reinterpret_cast<const float4*>(sbias2 + ((c0 + 4 * j) & cmask));
      r[4 * j] = __float_as_uint(q.x + bq.x); r[4 * j + 1] = __float_as_uint(q.y + bq.y);
      r[4 * j + 2] = __float_as_uint(q.z + bq.z); r[4 * j + 3] = __float_as_uint(q.w + bq.w);
    }
    tmem_st32(x_quad + (uint32_t)(32 * u), r);
  }
  tmem_st_wait();
}

// All MMAs of one conv, issued by the MMA warp (warp-uniform control flow, the instructions themselves predicated on the
// elected lane).  The issuing thread's own instruction stream sets the pace of these small MMAs (tools/micro/mma_chain.cu:
// ~30 scalar instructions between two MMAs cost 117 cycles against 48 of tensor work), so everything is hoisted: one
// instantiation per C (immediate strides), ring slots as precomputed descriptor words, a granule = several accumulators.
//   head group  [0, gh)          granule-outer, waits for the granule's e_done (plus the next one's when head_fwd)
//   middle      [gh, nts - gt)   stage-outer over all accumulators, after every e_done
//   tail group  [nts - gt, nts)  granule-outer, commits m_done per granule

template <int C, bool CG2>
__device__ __forceinline__ void resq_issue_conv(const ResGeom& g, bool leader, uint32_t w0_lo, uint64_t* b_full, uint64_t* b_empty,
                                                uint64_t* ew, uint64_t* mc, uint32_t par, uint32_t desc_hi, uint32_t a_tap0, uint32_t tap_step,
                                                uint32_t d_base, int& ib, uint32_t& pb, long long* wcyc) {
  const int ng = g.ng, nts = g.n_tstages, tb = g.tb, sbn = g.sb;
  const uint32_t stage_step = (uint32_t)g.bstage_bytes >> 4;
  const uint32_t gran_a = (uint32_t)g.gran * (uint32_t)(16 * C);      // descriptor step of one granule of S rows (128 rows x 2C bytes >> 4)
  const uint32_t gran_d = (uint32_t)(g.gran * C);                     // TMEM columns of one granule
  auto commit = [&](uint64_t* bar) { if (leader) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); } };
  int e_seen = 0;                                                     // granules [0, e_seen) of the input slab are ready
  auto wait_upto = [&](int gi) {
    if (gi > ng - 1) gi = ng - 1;
    if (e_seen > gi) return;
    const long long c0 = wcyc ? clock64() : 0;
    while (e_seen <= gi) { mbar_wait(&ew[e_seen], par); ++e_seen; }
    tc_fence_after();
    if (wcyc) wcyc[1] += clock64() - c0;
  };
  // ONE loop nest (one inlined copy of the MMA issue code per C: with a copy per group and stage the issuing warp's
  // instruction footprint grew several-fold and its issue rate fell to ~60 cycles per MMA): segment 0 = head group,
  // 1 = middle, 2 = tail group; the middle is a single pass over all accumulators.
  for (int seg = 0; seg < 3; ++seg) {
    const int s0 = seg == 0 ? 0 : (seg == 1 ? g.gh : nts - g.gt);
    const int s1 = seg == 0 ? g.gh : (seg == 1 ? nts - g.gt : nts);
    if (s0 >= s1) continue;
    const bool outer = seg != 1;
    const int passes = outer ? ng : 1;
    const int n_acc = outer ? g.gran : g.msub;
    const int fwd = seg == 0 ? g.head_fwd : 1;
    if (outer) {                                                        // the group's stages stay resident over all granules
      const long long c0 = wcyc ? clock64() : 0;
      int ibw = ib;
      uint32_t pbw = pb;
      for (int s = s0; s < s1; ++s) {
        mbar_wait(&b_full[ibw], pbw);
        if (++ibw == sbn) { ibw = 0; pbw ^= 1u; }
      }
      tc_fence_after();
      if (wcyc) wcyc[0] += clock64() - c0;
    }
    uint32_t a_g = a_tap0, d_g = d_base;
    for (int pass = 0; pass < passes; ++pass, a_g += gran_a, d_g += gran_d) {
      wait_upto(outer ? pass + fwd : ng - 1);
      int ibs = ib;
      uint32_t pbs = pb;
      for (int s = s0; s < s1; ++s) {
        if (!outer) {
          const long long c0 = wcyc ? clock64() : 0;
          mbar_wait(&b_full[ibs], pbs);
          tc_fence_after();
          if (wcyc) wcyc[0] += clock64() - c0;
        }
        res_issue_stage<C, CG2>(leader, n_acc, desc_hi, a_g, tap_step, w0_lo + (uint32_t)ibs * stage_step, s * tb, min(tb, g.k - s * tb), d_g);
        if (pass == passes - 1) commit(&b_empty[ibs]);                  // last user of the stage: hand the slot back right away
        if (++ibs == sbn) { ibs = 0; pbs ^= 1u; }
      }
      if (seg == 2) commit(&mc[pass]);
    }
    for (int s = s0; s < s1; ++s)
      if (++ib == sbn) { ib = 0; pb ^= 1u; }
  }
}

// One issuing warp per granule (g.iss2: two granules, warps 1 and 2 + ne).  With the phases overlapped, the epilogue warps of
// the issuing warp's scheduler compete with it for issue slots and a single issuer fell to 52-54 cycles per N = 64 MMA (43 as a
// CTA pair at the operand-fetch bound); with an issuer per granule the granule-outer walk disappears as well: each warp runs
// stage-outer over ITS accumulators, waits for its own granule (and the one before it: taps left of the centre) up front and
// for the next granule in front of the first stage that holds a tap right of the centre, and commits its granule's m_done.
template <int C, bool CG2>
__device__ __forceinline__ void resq_issue_conv_own(const ResGeom& g, int gran_idx, bool leader, uint32_t w0_lo, uint64_t* b_full,
                                                    uint64_t* b_empty, uint64_t* ew, uint64_t* mc, uint32_t par, uint32_t desc_hi, uint32_t a_tap0,
                                                    uint32_t tap_step, uint32_t d_base, int& ib, uint32_t& pb, long long* wcyc) {
  const int ng = g.ng, nts = g.n_tstages, tb = g.tb, sbn = g.sb;
  const uint32_t stage_step = (uint32_t)g.bstage_bytes >> 4;
  const uint32_t a_g = a_tap0 + (uint32_t)(gran_idx * g.gran) * (uint32_t)(16 * C);
  const uint32_t d_g = d_base + (uint32_t)(gran_idx * g.gran * C);
  auto commit = [&](uint64_t* bar) { if (leader) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); } };
  {
    const long long c0 = wcyc ? clock64() : 0;
    // (over its granules i, i + 2, ... an issuer waits for every e_done barrier at least once per conv: the parity waits stay meaningful)
    mbar_wait(&ew[gran_idx], par);
    if (gran_idx > 0) mbar_wait(&ew[gran_idx - 1], par);
    tc_fence_after();
    if (wcyc) wcyc[1] += clock64() - c0;
  }
  bool next_ready = gran_idx + 1 >= ng;
  const int first_right = ((g.k - 1) / 2 + 1) / tb;          // first stage with a tap right of the centre
  for (int ts = 0; ts < nts; ++ts) {
    {
      const long long c0 = wcyc ? clock64() : 0;
      mbar_wait(&b_full[ib], pb);
      if (wcyc) wcyc[0] += clock64() - c0;
    }
    if (!next_ready && ts >= first_right) {
      const long long c0 = wcyc ? clock64() : 0;
      mbar_wait(&ew[gran_idx + 1], par);
      next_ready = true;
      if (wcyc) wcyc[1] += clock64() - c0;
    }
    tc_fence_after();
    res_issue_stage<C, CG2>(leader, g.gran, desc_hi, a_g, tap_step, w0_lo + (uint32_t)ib * stage_step, ts * tb, min(tb, g.k - ts * tb), d_g);
    commit(&b_empty[ib]);
    if (++ib == sbn) { ib = 0; pb ^= 1u; }
  }
  if (!next_ready) { mbar_wait(&ew[gran_idx + 1], par); }     // k = 1: no tap right of the centre, the completion is still consumed
  commit(&mc[gran_idx]);
}

template <int MODE, bool CG2, bool WIDE>
__global__ void __maxnreg__(WIDE ? 96 : 168)
resq_tc_kernel(const __grid_constant__ ResMaps maps, const ResParams P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = P.c;
  const ResGeom& g = P.g;

  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slabA = smem;                                   // lrelu(x): input of every c1
  uint8_t* slabB = smem + (size_t)g.s_bytes;               // lrelu(c1 + b1): input of every c2
  uint8_t* stageB = slabB + (size_t)g.s_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageB + (size_t)g.sb * g.bstage_bytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + kResqMaxStages;
  uint64_t* m_done = b_empty + kResqMaxStages;             // [2][kResqMaxGran]: c1 / c2 commits per granule
  uint64_t* e_done = m_done + 2 * kResqMaxGran;            // [2][kResqMaxGran]: phase A / (phase B or slab-A load) per granule
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_done + 2 * kResqMaxGran);
  float* sbias = reinterpret_cast<float*>(bars + 72);      // [2 n_dil][C]: b1_s, b2_s
  float* epi_tiles = sbias + 2 * kResMaxDil * 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (P.span && threadIdx.x == 0) atomicMin(&P.span[0], (unsigned long long)gtime());

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2 * g.n_dil; ++i) tma_prefetch_desc(&maps.w[i]);
    for (int i = 0; i < g.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], g.iss2 ? 2u : 1u); }
    for (int i = 0; i < 2 * kResqMaxGran; ++i) {
      mbar_init(&m_done[i], 1);
      mbar_init(&e_done[i], (uint32_t)(CG2 ? 2 * g.ne : g.ne));
    }
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(tmem_slot, (uint32_t)g.tmem_cols); else tmem_alloc_dyn(tmem_slot, (uint32_t)g.tmem_cols); }
  if (warp >= 2) {
    // the pad rows above and below the tile are read by the outer taps but never written: zero them once, in both slabs
    const int pad_bytes = g.pad * g.rb;
    for (int sl = 0; sl < 2; ++sl) {
      uint8_t* lo = sl ? slabB : slabA;
      uint8_t* hi = lo + (size_t)(g.pad + g.mt) * g.rb;
      for (int o = (threadIdx.x - 64) * 16; o < pad_bytes; o += ((int)blockDim.x - 64) * 16) {
        *reinterpret_cast<uint4*>(lo + o) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(hi + o) = make_uint4(0u, 0u, 0u, 0u);
      }
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
  const int crank = CG2 ? (int)cluster_ctarank() : 0;
  const int walkers = CG2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int walk0 = CG2 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int walk_n = CG2 ? (g.total_items + 1) / 2 : g.total_items;
  auto item_of = [&](int wk) { return CG2 ? 2 * wk + crank : wk; };
  const int ng = g.ng;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (weights only), as in res_tc_kernel
    const bool leader = elect_one();
    int ib = 0;
    uint32_t pb = 0;
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      const int nxt = wk + walkers < walk_n ? item_of(wk + walkers) : g.total_items;
      if (leader && nxt < g.total_items) {      // the next item's x / branch-sum rows into L2, a whole item ahead
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
      const int rounds = g.iss2 ? g.ng / 2 : 1;       // an issuer per granule: the conv's stages once per round of two granules
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        for (int rd = 0; rd < rounds; ++rd)
        for (int ts = 0; ts < g.n_tstages; ++ts) {
          mbar_wait(&b_empty[ib], pb ^ 1u);
          if (leader) {
            if constexpr (CG2) {
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
    // -------------------------------------------------------------- MMA issuer(s): iss2 = one warp per granule
    const int my_gran = warp == 1 ? 0 : 1;
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template((uint32_t)g.rb);
    const uint32_t desc_hi = (uint32_t)(tmpl >> 32);
    const uint32_t desc_lo_fixed = (uint32_t)tmpl;
    const uint32_t row_step = (uint32_t)g.rb >> 4;
    const uint32_t sa_lo = desc_lo_fixed | ((smem_u32(slabA) & 0x3FFFFu) >> 4);
    const uint32_t sb_lo = desc_lo_fixed | ((smem_u32(slabB) & 0x3FFFFu) >> 4);
    const uint32_t w0_lo = desc_lo_fixed | ((smem_u32(stageB) & 0x3FFFFu) >> 4);     // weight ring slot i: + i * (bstage_bytes >> 4)
    int ib = 0;
    uint32_t pb = 0, n1 = 0;
    int ntr = warp == 1 ? 0 : 128;
    long long wcyc[2] = {0, 0};
    long long* wc = (P.trace && blockIdx.x == 0 && warp == 1) ? wcyc : nullptr;
    const long long cstart = clock64();
    for (int wk = (CG2 && crank != 0) ? walk_n : walk0; wk < walk_n; wk += walkers) {   // CTA pair: the leader issues for both
      for (int cv = 0; cv < 2 * g.n_dil; ++cv) {
        const int st = cv >> 1;
        const bool second = (cv & 1) != 0;
        const int dil = second ? 1 : g.dil[st];
        const int halo = second ? g.h2 : g.h1[st];
        uint64_t* ew = e_done + (second ? 0 : kResqMaxGran);      // c1 waits for phase B / the slab-A load, c2 for phase A
        uint64_t* mc = m_done + (second ? kResqMaxGran : 0);
        const uint32_t a_tap0 = (second ? sb_lo : sa_lo) + (uint32_t)(g.pad - halo) * row_step;
        const uint32_t d_base = second ? tmem_base : tmem_base + (uint32_t)acc_cols;   // c2 accumulates onto X, c1 onto its bias
        const uint32_t tap_step = (uint32_t)dil * row_step;
        L2S_RTRACE(128, ntr);
        if (g.iss2) {
          // ng / 2 rounds per conv: in round r this warp works on granule 2 r + my_gran, and the producer streams the conv's
          // weight stages once per round (a few KB each from L2): granules finish, and are rewritten by the epilogue warps,
          // round by round while the other round's MMAs run
          for (int gi = my_gran; gi < g.ng; gi += 2) {
            if (g.c == 64) resq_issue_conv_own<64, CG2>(g, gi, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
            else if (g.c == 32) resq_issue_conv_own<32, CG2>(g, gi, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
            else resq_issue_conv_own<16, CG2>(g, gi, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
          }
        } else
        if (g.c == 64) resq_issue_conv<64, CG2>(g, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
        else if (g.c == 32) resq_issue_conv<32, CG2>(g, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
        else resq_issue_conv<16, CG2>(g, leader, w0_lo, b_full, b_empty, ew, mc, n1 & 1u, desc_hi, a_tap0, tap_step, d_base, ib, pb, wc);
        L2S_RTRACE(128, ntr);
        if (second) ++n1;
      }
    }
    if (wc && lane == 0) { P.trace[500] = wcyc[0]; P.trace[501] = wcyc[1]; P.trace[502] = clock64() - cstart; }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    ResLane w;
    w.quad = warp & 3;
    w.half = (warp - 2) >> 2;
    w.ustep = g.ne >> 2;
    w.lane = lane;
    w.sw = g.rb == 128 ? (lane & 7) : (g.rb == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
    float* tile = epi_tiles + (size_t)(warp - 2) * g.tile_words;
    const uint32_t x_quad = tmem_base + ((uint32_t)(w.quad * 32) << 16);
    const uint32_t d1_quad = x_quad + (uint32_t)acc_cols;
    const int upg = (g.gran * g.c) >> 5;          // 32-column units per granule
    uint32_t n1 = 0;
    int ntr = warp == 2 ? 0 : 128;
    auto arrive = [&](uint64_t* bar) { if constexpr (CG2) mbar_arrive_cluster(bar, 0u, (uint32_t)crank); else mbar_arrive(bar); };   // pair: the leader's MMA thread waits
    auto publish = [&](uint64_t* bar) {          // S / X / D1 written by this warp are visible to the tensor core
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(bar);
    };
    auto publish_all = [&](uint64_t* base) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) for (int gr = 0; gr < ng; ++gr) arrive(&base[gr]);
    };
    struct Item { int b, q0, t_row0, lin, row_lim; };
    auto locate_item = [&](int wk) {
      Item it;
      const int item = item_of(wk);
      const bool dummy = item >= g.total_items;    // odd item count: the pair's last partner computes on zeros and stores nothing
      it.b = dummy ? 0 : item / g.m_items;
      const int mi = dummy ? 0 : item - it.b * g.m_items;
      it.q0 = mi * g.r_out;                        // first output row of the item
      it.t_row0 = it.q0 - g.h_tot;                 // position of tile row 0
      it.lin = dummy ? 0 : p.lin;
      it.row_lim = dummy ? 0 : min(p.lin, it.q0 + g.r_out);
      return it;
    };
    // phase after a conv, granule by granule.  EVERY warp waits for EVERY commit, also of granules it owns no unit of: a
    // parity wait is only meaningful for a warp that has seen the barrier's previous completion, and a warp that could
    // run ahead would arrive on an e_done barrier twice within one of its phases.
    auto phases = [&](uint64_t* md, uint64_t* ed, uint8_t* slab, uint32_t t_quad, const Item& it, bool edge, uint32_t par) {
      for (int gr = 0; gr < ng; ++gr) {
        const int lo = gr * upg, hi = lo + upg;
        mbar_wait(&md[gr], par);
        if (gr == 0) L2S_RTRACE(0, ntr);
        if (resq_first_unit(w, lo) < hi) {
          tc_fence_after();
          if (edge) resq_phase<true>(g, w, slab, t_quad, it.t_row0, it.lin, lo, hi);
          else resq_phase<false>(g, w, slab, t_quad, it.t_row0, it.lin, lo, hi);
        }
        publish(&ed[gr]);
      }
    };
    res_prebias_d1(g, w, d1_quad, sbias);            // bias of the first c1
    if (walk0 < walk_n) {
      const Item it = locate_item(walk0);
      res_load_x(P, w, slabA, x_quad, it.b, it.t_row0, sbias + 64, it.lin);
      publish_all(e_done + kResqMaxGran);
    }
    for (int wk = walk0; wk < walk_n; wk += walkers) {
      const Item it = locate_item(wk);
      const bool edge = it.t_row0 < 0 || it.t_row0 + g.mt > it.lin;     // some tile rows lie outside the utterance
      for (int st = 0; st < g.n_dil; ++st) {
        const uint32_t par = n1 & 1u;
        // ---- phase A: D1 -> slab B
        L2S_RTRACE(0, ntr);
        phases(m_done, e_done, slabB, d1_quad, it, edge, par);
        L2S_RTRACE(0, ntr);
        res_prebias_d1(g, w, d1_quad, sbias + (2 * ((st + 1) % g.n_dil)) * 64);   // while c2 runs
        if (st + 1 < g.n_dil) {
          // ---- phase B: X -> slab A
          L2S_RTRACE(0, ntr);
          phases(m_done + kResqMaxGran, e_done + kResqMaxGran, slabA, x_quad, it, edge, par);
          L2S_RTRACE(0, ntr);
          res_addbias_x(g, w, x_quad, sbias + (2 * (st + 1) + 1) * 64);           // while the next c1 runs
        } else {
          const bool has_next = wk + walkers < walk_n;
          Item nx{};
          if (has_next) {
            // ---- slab A of the next tile while the last c2 of this one runs (phase A above waited for the commit of
            //      the last granule: every c1 MMA of this step has retired and slab A has no reader left)
            nx = locate_item(wk + walkers);
            L2S_RTRACE(0, ntr);
            resq_load_s(P, w, slabA, nx.b, nx.t_row0, nx.lin);
            publish_all(e_done + kResqMaxGran);
            L2S_RTRACE(0, ntr);
          }
          // ---- output: X -> global (rows [q0, q0 + r_out) of the tile only)
          for (int gr = 0; gr < ng; ++gr) mbar_wait(&m_done[kResqMaxGran + gr], par);
          tc_fence_after();
          L2S_RTRACE(0, ntr);
          if constexpr (WIDE) {
            res_output<16, MODE, false>(p, tile, x_quad, it.b, it.t_row0, it.q0, it.row_lim, g.msub, g.c, w, nullptr, 0u);
          } else {
            if (g.c % 32 == 0) res_output<32, MODE, true>(p, tile, x_quad, it.b, it.t_row0, it.q0, it.row_lim, g.msub, g.c, w, nullptr, 0u);
            else res_output<16, MODE, true>(p, tile, x_quad, it.b, it.t_row0, it.q0, it.row_lim, g.msub, g.c, w, nullptr, 0u);
          }
          // ---- X of the next tile (the first phase A of the next tile, which releases its c2, comes after this in
          //      every warp's program order)
          L2S_RTRACE(0, ntr);
          if (has_next) resq_load_x(P, w, x_quad, nx.b, nx.t_row0, sbias + 64, nx.lin);
          L2S_RTRACE(0, ntr);
          tc_fence_before();
          __syncwarp();
        }
        ++n1;
      }
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

template <int MODE, bool CG2, bool WIDE>
inline cudaError_t launch_resq_mode(const ResParams& P, const ResMaps& maps, int grid, cudaStream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(resq_tc_kernel<MODE, CG2, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(resq_tc_kernel<MODE, CG2, WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
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
  return cudaLaunchKernelEx(&cfg, resq_tc_kernel<MODE, CG2, WIDE>, maps, P);
}

// Defined in tu_resq_tc.cu (the only translation unit that instantiates resq_tc_kernel).
#ifdef L2S_TU_RESQ_TC
cudaError_t launch_resq_tc(const ResParams& P, const ResMaps& maps, int grid, int mode, cudaStream_t stream) {
  const ResGeom& g = P.g;
  switch (mode) {
#define L2S_QMODE(m)                                                                                        \
  case m:                                                                                                   \
    if (g.ne == 16) return g.cg2 ? launch_resq_mode<m, true, true>(P, maps, grid, stream) : launch_resq_mode<m, false, true>(P, maps, grid, stream); \
    return g.cg2 ? launch_resq_mode<m, true, false>(P, maps, grid, stream) : launch_resq_mode<m, false, false>(P, maps, grid, stream);
    L2S_QMODE(4) L2S_QMODE(6) L2S_QMODE(8) L2S_QMODE(10) L2S_QMODE(12) L2S_QMODE(14)
#undef L2S_QMODE
    default: return cudaErrorInvalidValue;
  }
}
#endif  // L2S_TU_RESQ_TC

}  // namespace l2s
