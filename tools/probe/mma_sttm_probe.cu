// Probe: how fast does the tensor pipe execute tcgen05.mma.kind::i8 (M = 128, K = 32, A from TMEM, B from shared memory)
// at N = 64 and N = 32, alone and while expander-like warps stream tcgen05.st.32x32b.x16 into TMEM -- the two things
// sketch_i8_kernel does at the same time.
//   mma_sttm_probe [ctas_per_sm=2] [n_mma=4096] [N=64] [store_warps=8] [grid_sms=148] [unroll=1] [issuers=1] [commit_every=8] [store_delay=0]
// unroll = MMAs per elected region (1, 4 or 16); issuers = 1 (warp 1) or 2 (warps 1 and 3, n_mma each, own accumulators)
// Per resident CTA (same warp roles as the kernel): warp 1 issues n_mma MMAs back to back (a commit every 8) and waits
// for the last one; warps 4.. store 2 x 16 registers per iteration (4 KB per warp) + tcgen05.wait::st until the issuer
// is done.  Prints cycles per MMA (issue loop and completion) and cycles per store iteration, for CTA 0 and the mean.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../genomic_pca_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

struct Res {
  unsigned long long mma_issue, mma_done, st_cycles, st_iters;
};

template <int U>
__device__ __forceinline__ void issue_block(int it, uint32_t tb, uint32_t dcol, uint32_t bsm, uint32_t idesc) {
  const uint32_t desc_lo_const = (uint32_t)(1024 >> 4) << 16;
  const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t baddr = bsm + (uint32_t)(((it + u) & 7) * (32 * 64));
    const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo_const | ((baddr >> 4) & 0x3FFFu));
    tc_mma_ts_i8(tb + dcol, tb + 128 + 8 * ((it + u) & 15), bdesc, idesc, 1u);
  }
}

__global__ void __launch_bounds__(384, 2) probe(Res* out, int n_mma, int N, int store_warps, int tmem_cols, int spin,
                                                int unroll, int issuers, int commit_every, int store_delay) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const uint32_t base_s = smem_u32(sm);
  const uint32_t bsm = base_s;                // 16 KB operand stage (zeros)
  const uint32_t bar = base_s + 16384, slot = base_s + 16384 + 32;     // bar, bar + 8 (unwaited commits), bar + 24 (second issuer)
  volatile int* done = reinterpret_cast<volatile int*>(sm + 16384 + 40);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 24, 1);
    mbar_init(bar + 8, 1);
    fence_barrier_init();
    *done = 0;
  }
  if (warp == 1) {
    tmem_alloc(slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  tc_fence_after();
  uint32_t tb;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tb) : "r"(slot));
  const uint32_t A_COL0 = 128;
  if (warp == 1 || (warp == 3 && issuers == 2)) {
    const uint32_t dcol = warp == 1 ? 0u : 64u;
    const uint32_t barw = warp == 1 ? bar : bar + 24;
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const long long t0 = clock64();
    long long t1 = t0;
    if (n_mma > 0) {
      for (int it = 0; it < n_mma; it += unroll) {
        if (elect_one()) {
          if (unroll == 1) issue_block<1>(it, tb, dcol, bsm, idesc);
          else if (unroll == 4) issue_block<4>(it, tb, dcol, bsm, idesc);
          else issue_block<16>(it, tb, dcol, bsm, idesc);
          if (commit_every > 0 && ((it + unroll) % commit_every) == 0 && it + unroll < n_mma) tc_commit(bar + 8);     // (a commit nobody waits for: the kernel's tempty)
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) tc_commit(barw);
      __syncwarp();
      mbar_wait(barw, 0);
    } else {
      while (clock64() - t0 < spin) { }
      t1 = clock64();
    }
    const long long t2 = clock64();
    if (warp == 1) *done = 1;
    if (lane == 0 && warp == 1) {
      out[blockIdx.x].mma_issue = (unsigned long long)(t1 - t0);
      out[blockIdx.x].mma_done = (unsigned long long)(t2 - t0);
    }
  } else if (warp >= 4 && warp < 4 + store_warps) {
    const int quarter = warp & 3, tile = ((warp - 4) >> 2) & 1;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0x01010101u * (uint32_t)((lane + i) & 3);
    unsigned long long iters = 0;
    const long long t0 = clock64();
    while (!*done) {
      const uint32_t ta = tb + lane_addr + A_COL0 + ((iters & 1) * 2 + tile) * 32;
      tmem_st16(ta, r);
      tmem_st16(ta + 16, r);
      tc_wait_st();
      ++iters;
      if (store_delay > 0) {      // throttle: idle cycles between store iterations (sets the store rate)
        const long long d0 = clock64();
        while (clock64() - d0 < store_delay) { }
      }
    }
    const long long t1 = clock64();
    if (warp == 4 && lane == 0) {
      out[blockIdx.x].st_cycles = (unsigned long long)(t1 - t0);
      out[blockIdx.x].st_iters = iters;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tb, tmem_cols);
  }
}

int main(int argc, char** argv) {
  const int per_sm = argc > 1 ? atoi(argv[1]) : 2;
  const int n_mma = argc > 2 ? atoi(argv[2]) : 4096;
  const int N = argc > 3 ? atoi(argv[3]) : 64;
  const int store_warps = argc > 4 ? atoi(argv[4]) : 8;
  const int sms = argc > 5 ? atoi(argv[5]) : 148;
  const int unroll = argc > 6 ? atoi(argv[6]) : 1;
  const int issuers = argc > 7 ? atoi(argv[7]) : 1;
  const int commit_every = argc > 8 ? atoi(argv[8]) : 8;
  const int store_delay = argc > 9 ? atoi(argv[9]) : 0;
  const int grid = sms * per_sm;
  const int smem = per_sm == 2 ? 100 * 1024 : 200 * 1024;      // forces the residency asked for
  Res* d;
  cudaMalloc(&d, grid * sizeof(Res));
  cudaMemset(d, 0, grid * sizeof(Res));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe<<<grid, 384, smem>>>(d, n_mma, N, store_warps, per_sm == 2 ? 256 : 512, 400000, unroll, issuers, commit_every, store_delay);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("status %s\n", cudaGetErrorString(e));
    return 1;
  }
  Res* h = (Res*)malloc(grid * sizeof(Res));
  cudaMemcpy(h, d, grid * sizeof(Res), cudaMemcpyDeviceToHost);
  double issue = 0, done = 0, stc = 0, sti = 0;
  for (int b = 0; b < grid; ++b) {
    issue += (double)h[b].mma_issue;
    done += (double)h[b].mma_done;
    stc += (double)h[b].st_cycles;
    sti += (double)h[b].st_iters;
  }
  const int nm = n_mma > 0 ? n_mma : 1;
  printf("ctas/SM %d  N %d  mma %d x %d issuer(s), %d per elected region, commit every %d, store warps %d/CTA (+%d idle cycles per iteration): ",
         per_sm, N, n_mma, issuers, unroll, commit_every, store_warps, store_delay);
  if (n_mma > 0)
    printf("%.1f cycles per MMA per issuer (issue loop %.1f) = %.1f per MMA on the SM's pipe; ", done / grid / nm, issue / grid / nm,
           done / grid / nm / per_sm / issuers);
  if (store_warps > 0 && sti > 0)
    printf("%.1f cycles per store iteration (4 KB per warp) = %.1f B/clk/SM of TMEM stores", stc / sti,
           4096.0 * store_warps * per_sm / (stc / sti));
  printf("\n");
  return 0;
}
