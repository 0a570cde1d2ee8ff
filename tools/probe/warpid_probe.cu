// Probe: how are the warps of co-resident CTAs mapped to hardware warp slots (%warpid) and SMs?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 2) probe(unsigned* out, int spin) {
  extern __shared__ unsigned char sm[];
  unsigned wid, smid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 12 + warp) * 2 + 0] = wid;
    out[(blockIdx.x * 12 + warp) * 2 + 1] = smid;
  }
  // keep the CTA alive so that all CTAs are resident together
  long long t0 = clock64();
  while (clock64() - t0 < spin) sm[threadIdx.x] = (unsigned char)wid;
}
int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 384;
  const int grid = 296;
  unsigned* d;
  cudaMalloc(&d, grid * 12 * 2 * 4);
  cudaMemset(d, 0xff, grid * 12 * 2 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 115264);
  probe<<<grid, threads, 115264>>>(d, 2000000);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static unsigned h[296 * 12 * 2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int mism = 0, ctas_mism = 0;
  for (int b = 0; b < grid; ++b) {
    int any = 0;
    for (int w = 0; w < threads / 32; ++w)
      if ((h[(b * 12 + w) * 2] & 3) != (unsigned)(w & 3)) { ++mism; any = 1; }
    ctas_mism += any;
  }
  printf("threads %d: warps with (%%warpid & 3) != (warp & 3): %d, CTAs affected: %d of %d\n", threads, mism, ctas_mism, grid);
  for (int b = 0; b < 4; ++b) {
    printf("cta %d sm %u warpids:", b, h[(b * 12) * 2 + 1]);
    for (int w = 0; w < threads / 32; ++w) printf(" %u", h[(b * 12 + w) * 2]);
    printf("\n");
  }
  // find the partner on the same SM as cta 0
  for (int b = 1; b < grid; ++b)
    if (h[(b * 12) * 2 + 1] == h[1]) {
      printf("cta %d shares sm %u warpids:", b, h[1]);
      for (int w = 0; w < threads / 32; ++w) printf(" %u", h[(b * 12 + w) * 2]);
      printf("\n");
    }
  return 0;
}
