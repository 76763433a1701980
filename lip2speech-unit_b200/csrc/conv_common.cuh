// The one building block every layer of the generator maps onto: a
// "tap-offset convolution" over channels-last activations.
//
//   acc[b, q, n] = sum_j sum_ci in[b, q + tap_off[j], ci] * w[j][n][ci]        (rows outside [0, lin) read 0)
//   v            = (acc + bias[n] + res[idx] + acc_in[idx]) / div      (true division, as models.py:109)
//   out_raw[idx] = v                     (fp32, optional)
//   out_act[idx] = leaky_relu(v, slope)  (activation dtype, optional)
//   idx          = q * ntot + n + out_shift, written only when 0 <= idx < out_valid (per utterance)
//
// * Conv1d(C, C', k, dilation d)  (speech-resynthesis/models.py:20-31,78-79):
//     ntaps = k, tap_off[j] = (j - (k-1)/2) * d, ntot = C', out_shift = 0, mrows = L.
// * ConvTranspose1d(C, C', k, stride u, padding p) (models.py:82-86) in polyphase form:
//     output n = q*u + r - p gets taps j = r + m*u from input row q - m, so it is a
//     tap-offset conv with ntaps = ceil(k/u), tap_off[m] = -m, ntot = u*C'
//     (column r*C' + co), out_shift = -p*C', mrows = L + 1, and w[m][r*C'+co][ci] =
//     W[ci][co][r + m*u] (zero where r + m*u >= k).
// The leaky-ReLU that the reference applies to a conv's *input* (models.py:36-38,
// 101) is applied by the producer of that input, which is why every op can emit a
// raw fp32 copy (residual stream) and an activated copy (next conv's operand).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2s {

constexpr int kMaxTaps = 16;

struct ConvParams {
  const void* in;        // [B][lin][cin_pad]
  const void* w;         // [ntaps][ntot][cin_pad]
  const float* bias;     // [ntot]
  float* out_raw;        // may be null
  void* out_act;         // may be null
  const float* res;      // may be null
  const float* acc_in;   // may be null
  int batch, lin, cin_pad, ntaps, ntot, mrows;
  int tap_off[kMaxTaps];
  long long out_shift, out_valid;
  float div, slope;
  int act_f32;           // tcgen05 kernel only: 1 = operands / activated output are fp32 (tf32 mode), 0 = bf16
  int pf;               // prefetch experiment bits: 1 = L2 bulk prefetch of residual tiles, 2 = L1 prefetch of the next chunk
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// bf16 tensor-core mode: the activated copy is leaky_relu applied to the bf16-ROUNDED value, computed on
// packed pairs (cvt.rn.bf16x2 + HMUL2 + HMNMX2 = 3 instructions per 2 elements instead of 5).  It differs
// from bf16(leaky_relu(v)) by at most one bf16 ulp, on negative values only; every tensor-core epilogue
// (fused phase 1, fused phase 2, unfused convs) uses this one function, so they agree bit for bit.
__device__ __forceinline__ uint32_t lrelu_bf16x2(float a, float b, __nv_bfloat162 slope2) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 r = __hmax2(v, __hmul2(v, slope2));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// Finish W (4, 8 or 16) consecutive columns n0.. of row q for utterance b.
// idx granules are W-aligned (ntot, out_shift, out_valid are multiples of 16),
// so a granule is entirely inside or entirely outside the valid range.
template <typename Ta, int W>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, int b, int q, int n0, float (&v)[W]) {
  const long long idx = (long long)q * p.ntot + n0 + p.out_shift;
  if (idx < 0 || idx >= p.out_valid) return;
  const long long g = (long long)b * p.out_valid + idx;
#pragma unroll
  for (int i = 0; i < W; i += 4) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
    v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
  }
  if (p.res) {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      const float4 r = *reinterpret_cast<const float4*>(p.res + g + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (p.acc_in) {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      const float4 r = *reinterpret_cast<const float4*>(p.acc_in + g + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (p.div != 1.0f) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = __fdiv_rn(v[i], p.div);
  }
  if (p.out_raw) {
#pragma unroll
    for (int i = 0; i < W; i += 4)
      *reinterpret_cast<float4*>(p.out_raw + g + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (p.out_act) {
    if constexpr (sizeof(Ta) == 4) {
      float* o = reinterpret_cast<float*>(p.out_act) + g;
#pragma unroll
      for (int i = 0; i < W; i += 4)
        *reinterpret_cast<float4*>(o + i) = make_float4(lrelu(v[i], p.slope), lrelu(v[i + 1], p.slope),
                                                        lrelu(v[i + 2], p.slope), lrelu(v[i + 3], p.slope));
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out_act) + g;
      if constexpr (W % 8 == 0) {
#pragma unroll
        for (int i = 0; i < W; i += 8) {
          uint4 pk;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(lrelu(v[i], p.slope), lrelu(v[i + 1], p.slope));     pk.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(lrelu(v[i + 2], p.slope), lrelu(v[i + 3], p.slope)); pk.y = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(lrelu(v[i + 4], p.slope), lrelu(v[i + 5], p.slope)); pk.z = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(lrelu(v[i + 6], p.slope), lrelu(v[i + 7], p.slope)); pk.w = *reinterpret_cast<uint32_t*>(&t);
          *reinterpret_cast<uint4*>(o + i) = pk;
        }
      } else {
#pragma unroll
        for (int i = 0; i < W; i += 4) {
          uint2 pk;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(lrelu(v[i], p.slope), lrelu(v[i + 1], p.slope));     pk.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(lrelu(v[i + 2], p.slope), lrelu(v[i + 3], p.slope)); pk.y = *reinterpret_cast<uint32_t*>(&t);
          *reinterpret_cast<uint2*>(o + i) = pk;
        }
      }
    }
  }
}

}  // namespace l2s
