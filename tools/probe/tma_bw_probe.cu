// Probe: streaming bandwidth of 2-D TMA box loads from a [rows x pitch] byte matrix as a function of the box width,
// with the traversal order of the sketch kernel's sample-side pass (few rows, very long rows).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../genomic_pca_b200/csrc/tc_ptx.cuh"
using namespace tcptx;
struct P { uint32_t row_groups, stages_total, stages_per_split, n_items, box_w, ring, stage_bytes; };
__global__ void __launch_bounds__(128, 2) probe(const __grid_constant__ CUtensorMap tmap, const P p) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  const uint32_t bars = base + p.ring * p.stage_bytes;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.ring; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (16 + s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const uint32_t ks = item / p.row_groups, rg = item - ks * p.row_groups;
      const uint32_t st0 = ks * p.stages_per_split;
      uint32_t st1 = st0 + p.stages_per_split; if (st1 > p.stages_total) st1 = p.stages_total;
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const uint32_t s = it % p.ring, ph = (it / p.ring) & 1;
        mbar_wait(bars + 8 * (16 + s), ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bars + 8 * s, p.stage_bytes);
          tma_load_2d(base + s * p.stage_bytes, &tmap, bars + 8 * s, (int)(st * p.box_w), (int)(rg * 256));
          tma_load_2d(base + s * p.stage_bytes + p.stage_bytes / 2, &tmap, bars + 8 * s, (int)(st * p.box_w), (int)(rg * 256 + 128));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const uint32_t ks = item / p.row_groups;
      const uint32_t st0 = ks * p.stages_per_split;
      uint32_t st1 = st0 + p.stages_per_split; if (st1 > p.stages_total) st1 = p.stages_total;
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const uint32_t s = it % p.ring, ph = (it / p.ring) & 1;
        mbar_wait(bars + 8 * s, ph);
        if (elect_one()) mbar_arrive(bars + 8 * (16 + s));
        __syncwarp();
      }
    }
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? atoll(argv[1]) : 2504;
  const uint64_t pitch = argc > 2 ? atoll(argv[2]) : 2500096;
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  uint8_t* d; cudaMalloc(&d, rows * pitch); cudaMemset(d, 1, rows * pitch);
  const uint32_t row_groups = (uint32_t)((rows + 255) / 256);
  for (int l2p = 0; l2p < 2; ++l2p)
  for (uint32_t w : {64u, 128u, 256u}) {
    const uint32_t stage_bytes = 256 * w;
    const uint32_t ring = 65536 / stage_bytes;
    P p; p.row_groups = row_groups; p.box_w = w; p.stage_bytes = stage_bytes; p.ring = ring;
    p.stages_total = (uint32_t)(pitch / w);
    uint32_t ksplit = (4 * 296 + row_groups - 1) / row_groups;
    p.stages_per_split = (p.stages_total + ksplit - 1) / ksplit;
    ksplit = (p.stages_total + p.stages_per_split - 1) / p.stages_per_split;
    p.n_items = ksplit * row_groups;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {pitch, rows}; const cuuint64_t strides[1] = {pitch};
    const cuuint32_t box[2] = {w, 128}; const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     w == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     l2p ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed w=%u r=%d\n", w, (int)r); continue; }
    const int smem = ring * stage_bytes + 512;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      probe<<<296, 128, smem>>>(tm, p);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("rows %llu pitch %llu box %3u B x 128 rows, ring %u x %u KB, l2promo %d: %.3f ms  %.0f GB/s  (%s)\n",
           (unsigned long long)rows, (unsigned long long)pitch, w, ring, stage_bytes / 1024, l2p, best,
           (double)rows * pitch / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
