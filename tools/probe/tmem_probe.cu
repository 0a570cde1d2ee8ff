// Probe: TMEM allocation bases and store/load integrity for two co-resident CTAs.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../genomic_pca_b200/csrc/tc_ptx.cuh"
using namespace tcptx;
__global__ void __launch_bounds__(384, 2) probe(unsigned* out, unsigned* errs, int spin) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const uint32_t slot = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(base) : "r"(slot));
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) {
    out[blockIdx.x * 2] = base;
    out[blockIdx.x * 2 + 1] = smid;
  }
  if (warp >= 2 && warp < 10) {
    const int quarter = warp & 3, tile = (warp - 2) >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    for (int iter = 0; iter < 64; ++iter) {
      uint32_t r[32], v[32];
      for (int i = 0; i < 32; ++i) r[i] = (blockIdx.x << 20) ^ (iter << 12) ^ ((quarter * 32 + lane) << 5) ^ i ^ (tile << 30);
      tmem_st32(base + lane_addr + tile * 32, r);
      tc_wait_st();
      long long t0 = clock64();
      while (clock64() - t0 < spin) { }
      tmem_ld32(base + lane_addr + tile * 32, v);
      tc_wait_ld();
      int bad = 0;
      for (int i = 0; i < 32; ++i) bad += (v[i] != r[i]);
      if (bad) atomicAdd(errs, (unsigned)bad);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(base, 256);
  }
}
int main() {
  const int grid = 296;
  unsigned *d, *e;
  cudaMalloc(&d, grid * 2 * 4);
  cudaMalloc(&e, 4);
  cudaMemset(e, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 115264);
  probe<<<grid, 384, 115264>>>(d, e, 20000);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  static unsigned h[296 * 2];
  unsigned he = 0;
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(&he, e, 4, cudaMemcpyDeviceToHost);
  int n0 = 0, n256 = 0, other = 0;
  for (int b = 0; b < grid; ++b) {
    if (h[b * 2] == 0) ++n0; else if (h[b * 2] == 256) ++n256; else ++other;
  }
  printf("tmem bases: %d x 0, %d x 256, %d other; store/load mismatches: %u\n", n0, n256, other, he);
  for (int b = 0; b < grid; ++b)
    if (h[b * 2 + 1] == h[1]) printf("cta %d on sm %u base %u\n", b, h[b * 2 + 1], h[b * 2]);
  return 0;
}
