// Microbenchmark: issue rate of SS-mode tcgen05.mma (bf16, K = 16, M = 128 per CTA) as a function of N, with the
// operand access patterns of the time-packed ResBlock kernel (128-byte swizzled rows, row-shifted A descriptors,
// 32-byte K slices inside a row), with and without CTA pairs (cta_group::2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate tools/micro/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lip2speech-unit_b200/csrc/ptx.cuh"
using namespace l2s;

template <bool CG2>
__global__ void __launch_bounds__(128) mma_rate_kernel(int n, int iters, int pattern, int stream_w, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slab = smem;                       // 2 halves x 288 rows x 128 B
  uint8_t* wts = smem + 2 * 288 * 128;        // 8 groups x 128 rows x 128 B (CG2: 64 rows each)
  uint64_t* bar = reinterpret_cast<uint64_t*>(wts + 8 * 128 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (2 * 288 * 128 + 8 * 128 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) { if constexpr (CG2) tmem_alloc_cg2(slot, 512); else tmem_alloc_dyn(slot, 512); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int crank = CG2 ? (int)cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template(128);
    const uint32_t hi = (uint32_t)(tmpl >> 32), lo_fixed = (uint32_t)tmpl;
    const uint32_t s_lo = lo_fixed | ((smem_u32(slab) & 0x3FFFFu) >> 4);
    const uint32_t w_lo = lo_fixed | ((smem_u32(wts) & 0x3FFFFu) >> 4);
    const uint32_t idesc = umma_idesc_bf16(CG2 ? 256u : 128u, (uint32_t)n);
    t0 = clock64();
    if (crank == 0) {
      if (pattern == 0) {
        const uint64_t da = ((uint64_t)hi << 32) | (s_lo + (16u * 128u >> 4)), db = ((uint64_t)hi << 32) | w_lo;
        for (int i = 0; i < iters; ++i)
          if (leader) { if constexpr (CG2) umma_bf16_cg2(tmem + (uint32_t)((i & 1) * n), da, db, idesc, 1u); else umma_bf16(tmem + (uint32_t)((i & 1) * n), da, db, idesc, 1u); }
      } else {
        // packed-kernel pattern: 24 offsets = 6 weight groups x 4 K slices; A walks half select, row shift -1..+1, slice
        const uint32_t wstep = (uint32_t)(((CG2 ? 64 : 128) * 128) >> 4);
        for (int i = 0; i < iters; i += 24) {
#pragma unroll
          for (int o = 0; o < 24; ++o) {
            const uint32_t a = s_lo + (16u * 128u >> 4) + (uint32_t)(((o & 4) ? 288 * 128 : 0) >> 4) + (uint32_t)((((o >> 3) - 1) * 128) >> 4) + 2u * (uint32_t)(o & 3);
            const uint32_t b = w_lo + (uint32_t)(o >> 2) * wstep + 2u * (uint32_t)(o & 3);
            const uint64_t da = ((uint64_t)hi << 32) | a, db = ((uint64_t)hi << 32) | b;
            if (leader) { if constexpr (CG2) umma_bf16_cg2(tmem + (uint32_t)((o & 1) * n), da, db, idesc, 1u); else umma_bf16(tmem + (uint32_t)((o & 1) * n), da, db, idesc, 1u); }
          }
        }
      }
      if (leader) { if constexpr (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); }
    }
    mbar_wait(bar, 0);
    t1 = clock64();
  } else if (warp >= 2 && stream_w) {
    // competing shared-memory writes (what the weight TMA ring / epilogue stores do): plain 16-byte stores into the weight area tail
    uint4* dst = reinterpret_cast<uint4*>(wts + 6 * 128 * 128);
    for (int r = 0; r < stream_w; ++r)
      for (int i = threadIdx.x - 64; i < 2 * 128 * 128 / 16; i += 64) dst[i] = make_uint4(r, r, r, r);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG2) cluster_sync_all();
  if (warp == 1) { if constexpr (CG2) tmem_dealloc_cg2(tmem, 512); else tmem_dealloc_dyn(tmem, 512); }
  if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem = 2 * 288 * 128 + 8 * 128 * 128 + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4080;
  printf("cycles per K=16 MMA (M=128 per CTA), %d MMAs back to back, one CTA (pair) on an otherwise idle GPU\n", iters);
  printf("%6s %8s %10s %10s %10s %10s   tensor-bound cycles (N*128*16/4096 MAC/cyc... = N/2)\n", "N", "pattern", "1cta", "1cta+st", "cg2", "cg2+st");
  for (int pattern = 0; pattern < 2; ++pattern)
    for (int n : {16, 32, 64, 128, 256}) {
      double res[4];
      for (int v = 0; v < 4; ++v) {
        const bool cg2 = v >= 2;
        const int stream_w = (v & 1) ? 2500 : 0;
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          if (cg2) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, mma_rate_kernel<true>, n, iters, pattern, stream_w, d);
          } else {
            mma_rate_kernel<false><<<1, 128, smem>>>(n, iters, pattern, stream_w, d);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        res[v] = (double)h / iters;
      }
      printf("%6d %8d %10.1f %10.1f %10.1f %10.1f   %d\n", n, pattern, res[0], res[1], res[2], res[3], n / 2);
    }
  return 0;
}
