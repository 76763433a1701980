/*
 * l2s_debug.h -- test hooks and tuning knobs of libl2s_vocoder.so.  NOT part of the drop-in boundary
 * (include/l2s_vocoder.h): nothing a caller of the generator needs is declared here.  tests/, tools/ and
 * bench.py's per-launch timing use it.  l2s_debug_set takes the library's knob lock exclusively (every forward
 * holds it shared), so flipping a knob never races with a running forward; the knobs are process-wide.
 */
#ifndef L2S_DEBUG_H
#define L2S_DEBUG_H

#include "l2s_vocoder.h"

#ifdef __cplusplus
extern "C" {
#endif


/* Copy an intermediate buffer of the most recent forward to the host as fp32.
 * Names: "cond", "embed" (needs knob embed_tap), "conv_pre_act", and -- when the
 * forward was stopped with knob stop_after_stage = i -- "ups" (ups[i] output) and
 * "mrf" (stage i MRF mean).  Channels-last (B, L, C).  Synchronises. */
int l2s_debug_tap(l2s_vocoder* v, const char* name, float* host_dst, int64_t numel);

/* Run ONE generic tap-offset convolution (the building block every layer maps
 * onto) on caller-provided device buffers.  impl: 0 = CUDA-core kernel,
 * 1 = tcgen05 kernel (halo slab + row-shifted shared-memory descriptors),
 * 2 = retired probe (descriptor base-offset field filled in: measured WRONG on sm_100a,
 *     row-shifted descriptors need base_offset = 0),
 * 3 = one TMA-loaded A tile per tap, no row-shifted descriptors (probe / fallback).
 * `scale` is the divisor applied in the epilogue.  See csrc/conv_common.cuh. */
typedef struct l2s_conv_desc {
  const void* in;        /* [B][lin][cin_pad] channels-last, fp32 or bf16                */
  const void* w;         /* [ntaps][ntot][cin_pad] same dtype                            */
  const float* bias;     /* [ntot]                                                       */
  float* out_raw;        /* fp32, flat per-utterance index, may be NULL                  */
  void* out_act;         /* leaky-relu'd copy in the activation dtype, may be NULL       */
  const float* res;      /* fp32 residual, same indexing as out, may be NULL             */
  const float* acc_in;   /* fp32 running branch sum, may be NULL                         */
  int32_t act_bf16;      /* 1: in / w / out_act are bf16, 0: fp32                         */
  int32_t batch, lin, cin_pad, ntaps, ntot, mrows;
  int32_t tap_off[16];
  int64_t out_shift, out_valid;
  float scale, slope;
} l2s_conv_desc;
int l2s_debug_conv(const l2s_conv_desc* d, int32_t impl, int32_t device, void* stream, char* err, int32_t err_len);

/* With knob layer_events = 1 every launch of a forward is bracketed by a CUDA event
 * pair; this reads launch `idx` of the most recent forward (ms, algorithmic flops,
 * layer name).  L2S_ERR_INVALID past the last launch.  Synchronises on the event. */
int l2s_debug_layer_time(l2s_vocoder* v, int32_t idx, float* ms, double* flops, char* name, int32_t name_len);

/* Override a tuning / descriptor knob (tests and probes only): force_simt, stop_after_stage, stop_after_pre, per_tap,
 * sa_min, dual, cluster, cg2, alias_at, epi_tma, pdl, use_graph, fuse_pairs, fuse_branch, res_mode, res_msub,
 * res_single_pct, res_quad_pct, res_cg2, pack, pk_mode, pk_cg2, pk_single_pct, pk_fuse, pk_chan, pair_pref, pair_smem,
 * tc_cg2, epi_pf, trace_launch, span_ptr, plan_report, trace_ptr, max_msub, max_nt, slab_cap, max_ctas, embed_tap,
 * layer_events.  Waits for running forwards (exclusive knob lock). */
int l2s_debug_set(const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* L2S_DEBUG_H */
