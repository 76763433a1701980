// Probe: which (TMEM lane, column) does each register of tcgen05.ld.16x256b / 16x128b / 16x64b return?
// A warp writes lane l, column c = (l << 16) | c with the 32x32b shape, then reads back with the other shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_shapes tools/micro/tmem_shapes.cu && ./tmem_shapes
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lip2speech-unit_b200/csrc/ptx.cuh"
using namespace l2s;

__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int lane = threadIdx.x & 31;
  tmem_alloc_dyn(&slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = ((uint32_t)lane << 16) | (uint32_t)c;
  tmem_st32(base, v);
  for (int c = 0; c < 32; ++c) v[c] = ((uint32_t)lane << 16) | (uint32_t)(32 + c);
  tmem_st32(base + 32, v);
  tmem_st_wait();
  __syncwarp();
  uint32_t a[4], b[2], c1[1], d[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(c1[0]) : "r"(base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]) : "r"(base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  // second half of the quadrant: lane offset 16
  uint32_t e[4];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "r"(base + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out[lane * 32 + i] = a[i];
  for (int i = 0; i < 2; ++i) out[lane * 32 + 4 + i] = b[i];
  out[lane * 32 + 6] = c1[0];
  for (int i = 0; i < 8; ++i) out[lane * 32 + 8 + i] = d[i];
  for (int i = 0; i < 4; ++i) out[lane * 32 + 16 + i] = e[i];
  tc_fence_before();
  __syncthreads();
  tmem_dealloc_dyn(base, 64);
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 32 * 32 * 4);
  probe<<<1, 32>>>(d);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
  uint32_t h[32 * 32];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  auto show = [&](const char* name, int off, int n) {
    printf("%s: thread -> (lane,col) per register\n", name);
    for (int t = 0; t < 32; ++t) {
      printf("  t%2d:", t);
      for (int i = 0; i < n; ++i) printf(" (%2u,%2u)", h[t * 32 + off + i] >> 16, h[t * 32 + off + i] & 0xffff);
      printf("\n");
    }
  };
  show("16x256b.x1", 0, 4);
  show("16x128b.x1", 4, 2);
  show("16x64b.x1", 6, 1);
  show("16x256b.x2", 8, 8);
  show("16x256b.x1 at lane offset 16", 16, 4);
  return 0;
}
