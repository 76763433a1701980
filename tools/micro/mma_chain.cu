// Microbenchmark: what slows a stream of small tcgen05.mma instructions down?
//   chain  : consecutive K = 16 MMAs into the same TMEM tile before moving to the next of `nacc` tiles   (no effect, measured)
//   cevery : a tcgen05.commit onto a scratch mbarrier after every `cevery` MMAs (0 = only at the end)
//   side   : what four other warps (one per TMEM lane quadrant) do meanwhile: 0 nothing, 1 tcgen05.ld of other columns,
//            2 tcgen05.ld + tcgen05.st of other columns, 3 = 1 + fence.proxy.async + mbarrier arrive per load
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_chain tools/micro/mma_chain.cu && ./mma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lip2speech-unit_b200/csrc/ptx.cuh"
using namespace l2s;

__global__ void __launch_bounds__(192) mma_chain_kernel(int n, int iters, int chain, int nacc, int cevery, int side, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slab = smem;                       // 1056 rows x 128 B
  uint8_t* wts = smem + 1056 * 128;           // 8 taps x 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(wts + 8 * 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
  for (int i = threadIdx.x; i < (1056 * 128 + 8 * 64 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1u << 20); *stop = 0; fence_barrier_init(); }
  if (warp == 1) tmem_alloc_dyn(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t tmpl = umma_desc_template(128);
    const uint32_t hi = (uint32_t)(tmpl >> 32), lo_fixed = (uint32_t)tmpl;
    const uint32_t s_lo = lo_fixed | ((smem_u32(slab) & 0x3FFFFu) >> 4);
    const uint32_t w_lo = lo_fixed | ((smem_u32(wts) & 0x3FFFFu) >> 4);
    const uint32_t idesc = umma_idesc_bf16(128u, (uint32_t)n);
    t0 = clock64();
    int acc = 0, left = chain, tap = 0, cl = cevery;
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = s_lo + (uint32_t)(((16 + acc * 128 + (tap & 7) * 3) * 128) >> 4) + 2u * (uint32_t)(i & 3);
      const uint32_t b = w_lo + (uint32_t)(((tap & 7) * 64 * 128) >> 4) + 2u * (uint32_t)(i & 3);
      const uint64_t da = ((uint64_t)hi << 32) | a, db = ((uint64_t)hi << 32) | b;
      if (leader) umma_bf16(tmem + (uint32_t)(acc * n), da, db, idesc, 1u);
      if ((i & 3) == 3) ++tap;
      if (--left == 0) { left = chain; if (++acc == nacc) acc = 0; }
      if (cevery && --cl == 0) { cl = cevery; if (leader) umma_commit(&bar[1]); }
    }
    if (leader) umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    t1 = clock64();
    *stop = 1;
  } else if (warp >= 2 && side) {
    const uint32_t q = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    uint32_t r[32];
    uint32_t sink = 0;
    while (!*stop) {
      for (int u = 0; u < 8; ++u) {
        tmem_ld32(q + 32u * (uint32_t)(u & 3), r);
        tmem_ld_wait();
        sink += r[3];
        if (side == 2) { r[0] += 1u; tmem_st32(q + 32u * (uint32_t)(u & 3) + 128u, r); tmem_st_wait(); }
        if (side == 3) {
          *reinterpret_cast<uint4*>(slab + (size_t)(900 + (warp & 3) * 32 + (threadIdx.x & 31)) * 128 + 16 * (u & 7)) = make_uint4(r[0], r[1], r[2], r[3]);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if ((threadIdx.x & 31) == 0) mbar_arrive(&bar[2]);
        }
      }
    }
    if (sink == 0x12345u) out[1] = sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_dyn(tmem, 512);
  if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem = 1056 * 128 + 8 * 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("cycles per K=16 MMA (M=128, one CTA, idle GPU), %d MMAs, 4 accumulators, chains of 4\n", iters);
  const int cev[] = {0, 64, 32, 16, 8, 4};
  printf("%4s %5s  commit every:", "N", "side");
  for (int c : cev) printf(" %7d", c);
  printf("\n");
  for (int n : {16, 32, 64})
    for (int side = 0; side < 4; ++side) {
      printf("%4d %5d               ", n, side);
      for (int c : cev) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          mma_chain_kernel<<<1, 192, smem>>>(n, iters, 4, 4, c, side, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf(" %7.1f", (double)h / iters);
      }
      printf("\n");
    }
  return 0;
}
