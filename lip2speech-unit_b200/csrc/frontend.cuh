// Conditioning front end and waveform head: the HBM / latency bound ends of the
// path, written as plain coalesced CUDA-core kernels.
//
//   spk_project_kernel   spkr Linear(256,128)                 models_multi_input.py:37,80
//   cond_multi_kernel    dict gather -> ConvTranspose1d(E,E,4,2,1) -> exact GELU -> fc
//                        -> concat [mel | code feats | speaker] channels-last
//                                                             models_multi_input.py:67-82
//   cond_unit_kernel     unit-only parent: [dict[code] | spkr_table[id]]
//                                                             speech-resynthesis/models.py:188,214-217
//   post_kernel          leaky_relu(0.01) -> conv_post(k=7) -> tanh (+ int16)
//                                                             speech-resynthesis/models.py:110-112
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2s {

template <typename Ta>
__device__ __forceinline__ Ta to_act(float v);
template <>
__device__ __forceinline__ float to_act<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_act<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_mel(const void* mel, int dtype, long long i) {
  if (dtype == 0) return reinterpret_cast<const float*>(mel)[i];
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(mel)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(mel)[i]);
}

// out[b][c] = bias[c] + sum_k spkr[b][k] * W[c][k].  One block per utterance, one warp per output channel at a
// time: the lanes read a weight row contiguously and the partial sums are combined with shuffles (fixed order).
__global__ void spk_project_kernel(const float* __restrict__ spkr, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ out, int spk_dim, int e) {
  extern __shared__ float s_in[];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < spk_dim; k += blockDim.x) s_in[k] = spkr[(long long)b * spk_dim + k];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int c = warp; c < e; c += n_warps) {
    const float* wr = w + (long long)c * spk_dim;
    float acc = 0.f;
    for (int k = lane; k < spk_dim; k += 32) acc = fmaf(s_in[k], wr[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(long long)b * e + c] = acc + bias[c];
  }
}

struct CondParams {
  const long long* code;   // (B,U)
  const void* mel;         // (B,num_mels,T)
  int mel_dtype;
  const float* spk_vec;    // (B,E) projected speaker vectors (multi) or null
  const float* spk_raw;    // (B,spk_dim) raw speaker embeddings: when set, every block projects its utterance's embedding itself
  const float* spk_w;      //   (spkr Linear, [E][spk_dim] + spk_b) in the arithmetic order of spk_project_kernel, and spk_vec is
  const float* spk_b;      //   not read: one launch less in front of conv_pre
  int spk_dim;
  const float* dict;       // (num_embeddings,E)
  const float* tab;        // unit table folded through the ConvTranspose1d taps at load: [4][num_embeddings][E], tab[j][id] = W_j^T dict[id]
  const float* wt;         // unit ConvT weights repacked [4][E_in][E_out] (only used to build the table)
  const float* wt_bias;    // [E]
  const float* fc_t;       // fc weight transposed [E_in][E_out]
  const float* fc_bias;    // [E]
  void* cond;              // (B,T,cin_pad) channels-last, activation dtype
  float* embed_tap;        // optional (B,U,E): the raw gathered rows (bit-exactness test hook)
  int* err_flag;           // sticky out-of-range flag (host mapped)
  int batch, units, frames, e, num_mels, num_embeddings, cin_pad, has_spk;
};

constexpr int kCondFrames = 8;   // frames per pass (even start)
constexpr int kCondPasses = 3;   // passes per block: the in-block speaker projection (128 KB of weights from L2) is shared by 24 frames
                                 // (cfg2: 272 blocks = one wave at two 59-register blocks per SM; with 2 passes 400 blocks = two waves)
constexpr int kCondMaxSpk = 512; // largest spk_dim the in-block projection stages in shared memory
constexpr int kCondE = 128;      // embedding_dim this kernel is specialised for
constexpr int kCondSplit = 4;    // K (input channel) split: 4 thread groups of 128 each own a quarter of the reduction
constexpr int kCondThreads = kCondE * kCondSplit;

// One block = 8 frames x 128 output channels x 4 K-quarters.  With one thread per output channel (the first
// version) a forward put only ~20 warps on an SM and every warp walked 640 dependent weight loads: 82 us for cfg2,
// latency bound.  Splitting the reductions over four thread groups quadruples the warps for the same loads.
template <typename Ta>
__global__ void __launch_bounds__(kCondThreads) cond_multi_kernel(const CondParams p) {
  __shared__ float s_act[kCondFrames][kCondE];
  __shared__ float s_part[kCondSplit][kCondFrames][kCondE];
  __shared__ float s_sv[kCondE];
  __shared__ int s_id[kCondFrames / 2 + 2];
  const int b = blockIdx.y;
  const int c = threadIdx.x & (kCondE - 1);
  const int g = threadIdx.x >> 7;
  constexpr int KQ = kCondE / kCondSplit;                 // input channels per group
  constexpr int FPG = kCondFrames / kCondSplit;           // frames each group finishes
  Ta* cond = reinterpret_cast<Ta*>(p.cond) + ((long long)b * p.frames) * p.cin_pad;

  // (0) speaker projection of this utterance (models_multi_input.py:70-72), one warp per output channel at a time, lanes
  //     over the input channels, shuffle reduction: the arithmetic of spk_project_kernel, bit for bit
  if (p.has_spk) {
    if (p.spk_raw) {
      float* s_in = &s_part[0][0][0];                       // kCondMaxSpk floats fit the partial-sum buffer (not yet in use)
      for (int k = threadIdx.x; k < p.spk_dim; k += kCondThreads) s_in[k] = p.spk_raw[(long long)b * p.spk_dim + k];
      __syncthreads();
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      constexpr int NW = kCondThreads / 32, CPW = kCondE / NW;   // 16 warps, 8 channels each: all 8 dot products in flight at once
      float acc[CPW];
#pragma unroll
      for (int i = 0; i < CPW; ++i) acc[i] = 0.f;
      for (int k = lane; k < p.spk_dim; k += 32) {
        const float x = s_in[k];
#pragma unroll
        for (int i = 0; i < CPW; ++i) acc[i] = fmaf(x, p.spk_w[(long long)(warp + i * NW) * p.spk_dim + k], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < CPW; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (lane == 0) s_sv[warp + i * NW] = acc[i] + p.spk_b[warp + i * NW];
      }
    } else if (threadIdx.x < kCondE) {
      s_sv[threadIdx.x] = p.spk_vec[(long long)b * kCondE + threadIdx.x];
    }
  }
  __syncthreads();

  for (int pass = 0; pass < kCondPasses; ++pass) {
  const int t0 = (blockIdx.x * kCondPasses + pass) * kCondFrames;
  if (t0 >= p.frames) break;
  const int i0 = t0 >> 1;
  // (1) unit ids of this pass: i0-1 .. i0+4 (clamped, sticky error flag like the reference's device assert)
  if (threadIdx.x < kCondFrames / 2 + 2) {
    const int i = i0 - 1 + (int)threadIdx.x;
    int id = -1;                                            // -1: outside the utterance (contributes zero)
    if (i >= 0 && i < p.units) {
      long long v = p.code[(long long)b * p.units + i];
      if (v < 0 || v >= p.num_embeddings) {
        atomicOr(p.err_flag, 1);
        v = v < 0 ? 0 : p.num_embeddings - 1;
      }
      id = (int)v;
    }
    s_id[threadIdx.x] = id;
  }
  __syncthreads();
  if (p.embed_tap) {                                        // test hook: the raw gathered rows, an exact copy of the table
    for (int r = 1 + g; r <= kCondFrames / 2; r += kCondSplit) {
      const int i = i0 - 1 + r;
      if (i < p.units && s_id[r] >= 0) p.embed_tap[((long long)b * p.units + i) * kCondE + c] = p.dict[(long long)s_id[r] * kCondE + c];
    }
  }
  // (2) ConvTranspose1d(E,E,4,stride 2,pad 1) folded into the unit table: the layer acts on one of 200 table
  // rows, so W_j^T dict[id] is precomputed per (tap, id) at load and a frame is two gathers and an add:
  //   out[2i] = P1[u_i] + P3[u_{i-1}],  out[2i+1] = P2[u_i] + P0[u_{i+1}]   (+ bias)
  // (3) exact (erf) GELU.  Each thread group finishes two of the eight frames.
  {
    const float bias = p.wt_bias[c];
    const long long tap_stride = (long long)p.num_embeddings * kCondE;
#pragma unroll
    for (int j = 0; j < FPG; ++j) {
      const int f = g * FPG + j;
      const int h = f >> 1;                                 // unit i0 + h is s_id[h + 1]
      const int id0 = s_id[h + 1];
      const int idn = (f & 1) ? s_id[h + 2] : s_id[h];      // odd frames look ahead, even frames look back
      const int t_self = (f & 1) ? 2 : 1, t_nb = (f & 1) ? 0 : 3;
      float y = bias;
      if (id0 >= 0) y += p.tab[t_self * tap_stride + (long long)id0 * kCondE + c];
      if (idn >= 0) y += p.tab[t_nb * tap_stride + (long long)idn * kCondE + c];
      s_act[f][c] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f));
    }
  }
  __syncthreads();
  float acc[kCondFrames];
  // (4) fc, same split
#pragma unroll
  for (int f = 0; f < kCondFrames; ++f) acc[f] = 0.f;
#pragma unroll 8
  for (int k = g * KQ; k < (g + 1) * KQ; ++k) {
    const float w = p.fc_t[k * kCondE + c];
#pragma unroll
    for (int f = 0; f < kCondFrames; ++f) acc[f] = fmaf(s_act[f][k], w, acc[f]);
  }
#pragma unroll
  for (int f = 0; f < kCondFrames; ++f) s_part[g][f][c] = acc[f];
  __syncthreads();
  // (5) reduce + concat: [mel | code feats | speaker | zero pad]
  const float fb = p.fc_bias[c];
  const float sv = p.has_spk ? s_sv[c] : 0.f;
  const int spk_base = p.num_mels + kCondE;
#pragma unroll
  for (int j = 0; j < FPG; ++j) {
    const int f = g * FPG + j;
    const int t = t0 + f;
    if (t >= p.frames) break;
    const float y = fb + ((s_part[0][f][c] + s_part[1][f][c]) + (s_part[2][f][c] + s_part[3][f][c]));
    Ta* row = cond + (long long)t * p.cin_pad;
    row[p.num_mels + c] = to_act<Ta>(y);
    if (p.has_spk) row[spk_base + c] = to_act<Ta>(sv);
  }
  const int tail0 = spk_base + (p.has_spk ? kCondE : 0);
  for (int idx = threadIdx.x; idx < kCondFrames * (p.cin_pad - tail0); idx += kCondThreads) {
    const int f = idx / (p.cin_pad - tail0), ch = tail0 + idx % (p.cin_pad - tail0);
    if (t0 + f < p.frames) cond[(long long)(t0 + f) * p.cin_pad + ch] = to_act<Ta>(0.f);
  }
  for (int idx = threadIdx.x; idx < kCondFrames * p.num_mels; idx += kCondThreads) {
    const int m = idx / kCondFrames, f = idx % kCondFrames;
    const int t = t0 + f;
    if (t < p.frames)
      cond[(long long)t * p.cin_pad + m] =
          to_act<Ta>(load_mel(p.mel, p.mel_dtype, ((long long)b * p.num_mels + m) * p.frames + t));
  }
  __syncthreads();                                          // s_id / s_act / s_part are rewritten by the next pass
  }
}

// Unit-only variant: cond[b][u] = [dict[code[b][u]] | spk_table[id[b]] | 0 pad].
struct CondUnitParams {
  const long long* code;
  const long long* spk_id;     // (B) or null
  const float* dict;
  const float* spk_table;
  void* cond;
  float* embed_tap;
  int* err_flag;
  int batch, units, e, num_embeddings, num_speakers, cin_pad, has_spk;
};

template <typename Ta>
__global__ void cond_unit_kernel(const CondUnitParams p) {
  const int b = blockIdx.y;
  const int u = blockIdx.x;
  Ta* row = reinterpret_cast<Ta*>(p.cond) + ((long long)b * p.units + u) * p.cin_pad;
  long long id = p.code[(long long)b * p.units + u];
  if (id < 0 || id >= p.num_embeddings) {
    if (threadIdx.x == 0) atomicOr(p.err_flag, 1);
    id = id < 0 ? 0 : p.num_embeddings - 1;
  }
  long long sid = 0;
  if (p.has_spk) {
    sid = p.spk_id[b];
    if (sid < 0 || sid >= p.num_speakers) {
      if (threadIdx.x == 0) atomicOr(p.err_flag, 2);
      sid = sid < 0 ? 0 : p.num_speakers - 1;
    }
  }
  for (int c = threadIdx.x; c < p.cin_pad; c += blockDim.x) {
    float v = 0.f;
    if (c < p.e) {
      v = p.dict[id * p.e + c];
      if (p.embed_tap) p.embed_tap[((long long)b * p.units + u) * p.e + c] = v;
    } else if (p.has_spk && c < 2 * p.e) {
      v = p.spk_table[sid * p.e + (c - p.e)];
    }
    row[c] = to_act<Ta>(v);
  }
}

// Waveform head.  in: fp32 channels-last (B,L,C) MRF mean of the last stage.
constexpr int kPostMaxC = 64;

struct PostParams {
  const float* in;
  float wc[7 * kPostMaxC];   // [7][C] weights, passed by value: they sit in the constant bank and feed the FMAs directly
                             // (read from shared memory they doubled the kernel's shared-memory traffic)
  float bias;
  float* out;          // (B,L) or null
  int16_t* out_i16;    // (B,L) or null
  int batch, len, c;
};

constexpr int kPostTile = 256;

// pitch (floats) of a staged row: a multiple of 4 (float4 reads) that spreads 8 consecutive rows over all banks
__host__ __device__ inline int post_pitch(int c) { return c + 4; }

// CC > 0: the channel count as a compile-time constant (the index arithmetic of the staging loop and the tap loops
// unroll: the run-time version spent two thirds of its 22.6 M warp instructions on divisions and loop control and was
// issue-bound at 31 us for cfg2; 0 = any C <= kPostMaxC.
template <int CC>
__global__ void __launch_bounds__(kPostTile) post_kernel(const PostParams p) {
  extern __shared__ __align__(16) float s_x[];   // [(kPostTile + 6)][post_pitch(C)]
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kPostTile;
  const int c = CC > 0 ? CC : p.c, pitch = post_pitch(c), c4 = c >> 2;
  const float4* in4 = reinterpret_cast<const float4*>(p.in + (long long)b * p.len * c);
  const int n4 = (kPostTile + 6) * c4;
#pragma unroll 2
  for (int i = threadIdx.x; i < n4; i += kPostTile) {
    const int r = i / c4, q = i - r * c4;
    const int l = l0 - 3 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l >= 0 && l < p.len) {
      v = in4[(long long)l * c4 + q];
      // F.leaky_relu default slope 0.01, models.py:110
      v.x = v.x > 0.f ? v.x : v.x * 0.01f; v.y = v.y > 0.f ? v.y : v.y * 0.01f;
      v.z = v.z > 0.f ? v.z : v.z * 0.01f; v.w = v.w > 0.f ? v.w : v.w * 0.01f;
    }
    *reinterpret_cast<float4*>(s_x + r * pitch + q * 4) = v;
  }
  __syncthreads();
  const int l = l0 + threadIdx.x;
  if (l >= p.len) return;
  float acc = p.bias;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float4* xr = reinterpret_cast<const float4*>(s_x + (threadIdx.x + j) * pitch);
    const float* wr = p.wc + j * c;
#pragma unroll
    for (int q = 0; q < (CC > 0 ? CC / 4 : 1); ++q) {
      if (CC > 0) {
        const float4 x = xr[q];
        acc = fmaf(x.x, wr[4 * q], acc); acc = fmaf(x.y, wr[4 * q + 1], acc);
        acc = fmaf(x.z, wr[4 * q + 2], acc); acc = fmaf(x.w, wr[4 * q + 3], acc);
      }
    }
    if (CC == 0) {
      for (int q = 0; q < c4; ++q) {
        const float4 x = xr[q];
        acc = fmaf(x.x, wr[4 * q], acc); acc = fmaf(x.y, wr[4 * q + 1], acc);
        acc = fmaf(x.z, wr[4 * q + 2], acc); acc = fmaf(x.w, wr[4 * q + 3], acc);
      }
    }
  }
  const float y = tanhf(acc);
  const long long o = (long long)b * p.len + l;
  if (p.out) p.out[o] = y;
  if (p.out_i16) {
    // inference.py:79-81: audio * 32768 -> int16 (astype truncates toward zero); saturate instead of wrapping
    float s = y * 32768.0f;
    s = fminf(fmaxf(s, -32768.0f), 32767.0f);
    p.out_i16[o] = (int16_t)s;
  }
}

// Row-per-thread head (C = 16 / 32): every thread reads ITS row once (64 / 128 contiguous bytes straight from global, no
// staging), forms the row's seven per-tap dot products p_j = sum_c w[j][c] * lrelu(x[l][c]), and an output sample is
// bias + sum_j p_j[o - 3 + j], collected from the neighbouring rows through 7 KB of shared memory.  post_kernel stages a
// tile and every output re-reads seven rows of it (7x the tile through the shared-memory pipe: 27-31 us for cfg2, 0.32-0.35
// of the copy bandwidth); this one is bound by the 65 MB read.  The sum is formed per tap first, then over the taps (the
// reference's own cuDNN / MKL order is unspecified; post_kernel's single chain differs in the last fp32 bits).
constexpr int kPostRowsOut = kPostTile - 6;   // output samples per block

template <int CC>
__global__ void __launch_bounds__(kPostTile) post_rows_kernel(const PostParams p) {
  __shared__ float sp[7][kPostTile + 8];
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kPostRowsOut;
  const int r = threadIdx.x;                  // block row r is sample l0 - 3 + r
  const int l = l0 - 3 + r;
  float part[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) part[j] = 0.f;
  if (l >= 0 && l < p.len) {                  // rows outside the utterance are conv_post's zero padding
    const float4* in4 = reinterpret_cast<const float4*>(p.in + ((long long)b * p.len + l) * CC);
    float4 v[CC / 4];
#pragma unroll
    for (int q = 0; q < CC / 4; ++q) v[q] = __ldg(in4 + q);
#pragma unroll
    for (int q = 0; q < CC / 4; ++q) {
      // F.leaky_relu default slope 0.01, models.py:110
      const float x0 = v[q].x > 0.f ? v[q].x : v[q].x * 0.01f, x1 = v[q].y > 0.f ? v[q].y : v[q].y * 0.01f;
      const float x2 = v[q].z > 0.f ? v[q].z : v[q].z * 0.01f, x3 = v[q].w > 0.f ? v[q].w : v[q].w * 0.01f;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float* wr = p.wc + j * CC + 4 * q;
        part[j] = fmaf(x0, wr[0], part[j]); part[j] = fmaf(x1, wr[1], part[j]);
        part[j] = fmaf(x2, wr[2], part[j]); part[j] = fmaf(x3, wr[3], part[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) sp[j][r] = part[j];
  __syncthreads();
  const int o = l0 + r;
  if (r >= kPostRowsOut || o >= p.len) return;
  float acc = p.bias;
#pragma unroll
  for (int j = 0; j < 7; ++j) acc += sp[j][r + j];   // tap j reads sample o - 3 + j = block row r + j
  const float y = tanhf(acc);
  const long long oo = (long long)b * p.len + o;
  if (p.out) p.out[oo] = y;
  if (p.out_i16) {
    // inference.py:79-81: audio * 32768 -> int16 (astype truncates toward zero); saturate instead of wrapping
    float s = y * 32768.0f;
    s = fminf(fmaxf(s, -32768.0f), 32767.0f);
    p.out_i16[oo] = (int16_t)s;
  }
}

}  // namespace l2s
