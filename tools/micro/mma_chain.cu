// Microbenchmarks behind two scheduling decisions of the whole-ResBlock kernels (res_tc.cuh / resq_tc.cuh):
//  (1) ORDER: `chain` consecutive K = 16 MMAs accumulate into the same TMEM tile before the issuer moves to the next of `nacc`
//      tiles.  Measured: no effect (48.1 cycles per N = 64 MMA for every chain length 1..64).
//  (2) OVERLAP: what a stream of small MMAs loses when four other warps (one per TMEM lane quadrant) do epilogue-like work
//      at the same time: side 0 nothing, 1 tcgen05.ld of other columns, 2 tcgen05.ld + 16-byte shared-memory stores
//      (a phase A / B), 3 global loads + shared-memory stores (the next tile's load), 4 = 2 + fence.proxy.async per unit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_chain tools/micro/mma_chain.cu && ./mma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lip2speech-unit_b200/csrc/ptx.cuh"
using namespace l2s;

__global__ void __launch_bounds__(192) mma_side_kernel(int n, int iters, int side, const float4* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slab = smem;                       // 1056 rows x 128 B
  uint8_t* wts = smem + 1056 * 128;           // 8 taps x 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(wts + 8 * 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
  for (int i = threadIdx.x; i < (1056 * 128 + 8 * 64 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); *stop = 0; fence_barrier_init(); }
  if (warp == 1) tmem_alloc_dyn(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0, units = 0;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template(128);
    const uint32_t hi = (uint32_t)(tmpl >> 32), lo_fixed = (uint32_t)tmpl;
    const uint32_t s_lo = lo_fixed | ((smem_u32(slab) & 0x3FFFFu) >> 4);
    const uint32_t w_lo = lo_fixed | ((smem_u32(wts) & 0x3FFFFu) >> 4);
    const uint32_t idesc = umma_idesc_bf16(128u, (uint32_t)n);
    const uint64_t da = ((uint64_t)hi << 32) | (s_lo + (16u * 128u >> 4)), db = ((uint64_t)hi << 32) | w_lo;
    t0 = clock64();
    for (int i = 0; i < iters; ++i)
      if (leader) umma_bf16(tmem + (uint32_t)((i & 1) * n), da, db, idesc, 1u);
    if (leader) umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    t1 = clock64();
    *stop = 1;
  } else if (warp >= 2 && side) {
    const uint32_t q = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    uint8_t* my = slab + (size_t)(600 + (warp & 3) * 32 + lane) * 128;     // rows the MMAs do not read
    uint32_t r[32];
    uint32_t sink = 0;
    while (!*stop) {
      for (int u = 0; u < 4; ++u) {
        if (side == 3) {
          const float4* src = gsrc + ((size_t)(units & 1023) * 128 + (warp & 3) * 32 + lane) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 v = __ldg(src + j); r[4 * j] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y); r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w); }
        } else {
          tmem_ld32(q + 32u * (uint32_t)u, r);
          tmem_ld_wait();
        }
        sink += r[3];
        if (side >= 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e)   // 32 fp32 -> 16 bf16 pairs -> four 16-byte slots (swizzled like the S slab)
            *reinterpret_cast<uint4*>(my + ((e ^ (lane & 7)) << 4)) = make_uint4(r[8 * e] ^ r[8 * e + 1], r[8 * e + 2] ^ r[8 * e + 3], r[8 * e + 4] ^ r[8 * e + 5], r[8 * e + 6] ^ r[8 * e + 7]);
          if (side == 4) { fence_proxy_async_smem(); tc_fence_before(); __syncwarp(); }
        }
        ++units;
      }
    }
    if (sink == 0x12345u) out[7] = sink;
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&out[1]), (unsigned long long)units);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem, 512);
  if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  float4* g;
  cudaMalloc(&g, (size_t)1024 * 128 * 8 * sizeof(float4));
  cudaMemset(g, 0, (size_t)1024 * 128 * 8 * sizeof(float4));
  const int smem = 1056 * 128 + 8 * 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(mma_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 8192;
  printf("cycles per K=16 MMA (M=128, one CTA, idle GPU), %d MMAs back to back; side work of 4 warps in 32x32 units (see header)\n", iters);
  printf("%4s %5s %12s %22s\n", "N", "side", "cyc/MMA", "side units per 1000 cyc");
  for (int n : {16, 32, 64, 128})
    for (int side = 0; side < 5; ++side) {
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d, 0, 64);
        mma_side_kernel<<<1, 192, smem>>>(n, iters, side, g, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      }
      printf("%4d %5d %12.1f %22.2f\n", n, side, (double)h[0] / iters, 1000.0 * (double)h[1] / (double)h[0]);
    }
  return 0;
}
