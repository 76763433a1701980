// CUDA-core implementation of the tap-offset convolution (conv_common.cuh).
// It is the fp32 precision mode and the on-device cross-check for the tcgen05
// kernel (it accepts bf16 operands too, with the same fp32 accumulation).
// 64 x 64 output tile per 256-thread block, 4 x 4 outputs per thread, 16-channel
// K slices staged (transposed) through shared memory.
#pragma once
#include "conv_common.cuh"

namespace l2s {

template <typename Ta>
__device__ __forceinline__ void load4(const Ta* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

constexpr int kSimtTile = 64;
constexpr int kSimtK = 16;
constexpr int kSimtPitch = kSimtTile + 4;

template <typename Ta>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
  __shared__ __align__(16) float As[kSimtK][kSimtPitch];
  __shared__ __align__(16) float Ws[kSimtK][kSimtPitch];

  const int b = blockIdx.z;
  const int q0 = blockIdx.x * kSimtTile;
  const int n0 = blockIdx.y * kSimtTile;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int lrow = tid >> 2;        // 0..63: tile row (A) / tile column (W) this thread stages
  const int lk = (tid & 3) * 4;     // 4 consecutive channels of the 16-channel slice

  const Ta* in = reinterpret_cast<const Ta*>(p.in) + (long long)b * p.lin * p.cin_pad;
  const Ta* w = reinterpret_cast<const Ta*>(p.w);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int j = 0; j < p.ntaps; ++j) {
    const int row = q0 + lrow + p.tap_off[j];
    const bool row_ok = row >= 0 && row < p.lin;
    const bool col_ok = (n0 + lrow) < p.ntot;
    const Ta* arow = in + (long long)row * p.cin_pad;
    const Ta* wrow = w + ((long long)j * p.ntot + n0 + lrow) * p.cin_pad;
    for (int c0 = 0; c0 < p.cin_pad; c0 += kSimtK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok) load4<Ta>(arow + c0 + lk, av);
      if (col_ok) load4<Ta>(wrow + c0 + lk, wv);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[lk + i][lrow] = av[i];
        Ws[lk + i][lrow] = wv[i];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kSimtK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 ww = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w};
        const float bb[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(aa[i], bb[jj], acc[i][jj]);
      }
    }
  }

  const int n = n0 + tx * 4;
  if (n >= p.ntot) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= p.mrows) continue;
    conv_epilogue<Ta, 4>(p, b, q, n, acc[i]);
  }
}

template <typename Ta>
inline cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t stream) {
  dim3 grid((p.mrows + kSimtTile - 1) / kSimtTile, (p.ntot + kSimtTile - 1) / kSimtTile, p.batch);
  conv_simt_kernel<Ta><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace l2s
