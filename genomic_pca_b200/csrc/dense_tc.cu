// dense_tc.cu -- tcgen05 engine for the skinny products with the DENSE fp32 condensed-feature matrix of EigenSNP's
// global randomized SVD (the stage configured at src/main.rs:317-318, run inside EigenSNPCoreAlgorithm::compute_pca,
// src/main.rs:365).  Replaces the cuBLAS SGEMM calls of round 1.
//
//   out[r, :] = a_r * sum_k X[r, k] * f_k * B[k, :]  -  b_r * sum_k e_k * B[k, :]          (l <= 32 columns)
//
// with X a view of the row-major matrix C [N x R] (row stride ldc floats):
//   mode ROWS : X = C       rows r = samples,            k = condensed features (contiguous in memory)
//   mode COLS : X = C^T     rows r = condensed features, k = samples            (a strided column of C per row)
// -- the same form as a sketch pass, so the column standardisation of the condensed matrix (z = (c - mean) / sd) is
// folded into f / e / a / b and the standardised matrix is never written: the round-1 code swept C three more times.
//
// Arithmetic: split-bf16.  Every fp32 value is the sum of two bf16 numbers hi + lo up to 2^-17 relative; the tensor
// cores form  Xh*Bh + Xh*Bl + Xl*Bh  with fp32 accumulation in TMEM (the dropped Xl*Bl term is 2^-16 of a product).
//   A = X tile, converted in registers (fp32 -> bf16 hi | lo) by the thread that owns the output row and stored to TMEM
//       (tcgen05.st) -- the expanded operand never touches shared memory, as in the genotype engines;
//   B = [Bh | Bl] image (UMMA K-major core matrices, bf16) streamed by 1-D bulk copies;
//   per 16 k: one MMA  Xh x [Bh|Bl] (N = 64: two accumulators)  and one  Xl x Bh (N = 32, into the first).
// The pass is bound by the HBM stream of C (4 bytes per element, a handful of instructions per element on the
// expanders).  Same roles / rings / persistent round-robin schedule as sketch_i8.cu.
#include <cuda.h>
#include <cuda_bf16.h>

#include "dense_tc.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int RT = 2;              // row tiles of 128 rows per CTA
constexpr int KS = 32;             // k per pipeline stage (128 bytes of a row in ROWS mode)
constexpr int NL = 32;             // logical columns
constexpr int NC = 2 * NL;         // image columns: hi | lo
constexpr int NUM_THREADS = 384;   // warp 0 A producer, 1 MMA issuer, 2 B producer, 3 idle, 4..11 expanders
constexpr int SA = 3, SB = 3, SLOTS = 2;
constexpr int A_TILE_BYTES = 128 * KS * 4;              // 16 KB
constexpr int A_STAGE_BYTES = RT * A_TILE_BYTES;        // 32 KB
constexpr int B_STAGE_BYTES = KS * NC * 2;              // 4 KB
constexpr int TMEM_COLS = 256;
constexpr int D_COL0 = 0;                               // RT * NC accumulator columns
constexpr int A_COL0 = RT * NC;                         // SLOTS * RT * 32 columns (16 hi + 16 lo per tile)
static_assert(RT * NC + SLOTS * RT * 32 <= TMEM_COLS, "TMEM budget");
constexpr int SMEM_BYTES = SA * A_STAGE_BYTES + SB * B_STAGE_BYTES + 256 + 320;
// registers move from the producer / issuer warpgroup to the expanders (setmaxnreg works per warpgroup)
#define DENSE_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 32;")
#define DENSE_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 104;")

struct DenseParams {
  const __nv_bfloat16* bimg;   // [total_stages][KS * NC]
  uint64_t rows;
  uint32_t total_stages, stages_per_split, ksplit, row_groups, n_items;
  const float* a;
  const float* b;
  const float* cvec;           // [32]
  float* out;
  uint32_t ldo, l;
  float* partial;              // [ksplit][rows][32]
};

// two fp32 -> (bf16x2 hi, bf16x2 lo), element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16);
  const float r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 q = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&q);
}

template <bool COLS>
__global__ void __launch_bounds__(NUM_THREADS, 2) dense_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const DenseParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = smem_base + SA * A_STAGE_BYTES;
  const uint32_t bars = b_ring + SB * B_STAGE_BYTES;
  auto bar_afull = [&](int s) { return bars + 8u * s; };
  auto bar_aempty = [&](int s) { return bars + 8u * (4 + s); };
  auto bar_bfull = [&](int s) { return bars + 8u * (8 + s); };
  auto bar_tfull = [&](int j) { return bars + 8u * (12 + j); };
  auto bar_tempty = [&](int j) { return bars + 8u * (20 + j); };
  auto bar_bempty = [&](int s) { return bars + 8u * (24 + s); };
  const uint32_t bar_accfull = bars + 8u * 28;
  const uint32_t bar_accempty = bars + 8u * 29;
  const uint32_t tmem_slot = bars + 8u * 30;
  float* cv_s = reinterpret_cast<float*>(smem_raw + (bars + 256 - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(bar_afull(s), 1);
      mbar_init(bar_aempty(s), 4 * RT);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_bempty(s), 1);
    }
    for (int j = 0; j < SLOTS; ++j) {
      mbar_init(bar_tfull(j), 4 * RT);
      mbar_init(bar_tempty(j), 1);
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 4 * RT);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NL) cv_s[threadIdx.x - 64] = p.cvec[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto item_range = [&](uint32_t item, uint32_t& rg, uint32_t& ks, uint32_t& st0, uint32_t& st1) {
    ks = item / p.row_groups;
    rg = item - ks * p.row_groups;
    st0 = ks * p.stages_per_split;
    st1 = st0 + p.stages_per_split;
    if (st1 > p.total_stages) st1 = p.total_stages;
  };

  if (warp == 0 || warp == 2) {
    DENSE_REG_DEC();
    // one elected thread runs the whole role (an elected region per stage costs ~100 cycles: tools/probe/mma_sttm_probe.cu)
    const bool is_a = (warp == 0);
    uint32_t it = 0;
    const bool leader = elect_one();
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x) {
      uint32_t rg, ks, st0, st1;
      item_range(item, rg, ks, st0, st1);
      const int row0 = (int)(rg * (RT * 128));
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        if (is_a) {
          const int s = it % SA;
          mbar_wait(bar_aempty(s), ((it / SA) & 1u) ^ 1u);
          const uint32_t sbase = a_ring + s * A_STAGE_BYTES;
          mbar_arrive_expect_tx(bar_afull(s), A_STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < RT; ++t) {
            if (COLS)      // box [128 rows-of-X (inner, contiguous in C) x 32 k]
              tma_load_2d(sbase + t * A_TILE_BYTES, &tmap, bar_afull(s), row0 + t * 128, (int)(st * KS));
            else           // box [32 k (inner) x 128 rows]
              tma_load_2d(sbase + t * A_TILE_BYTES, &tmap, bar_afull(s), (int)(st * KS), row0 + t * 128);
          }
        } else {
          const int s = it % SB;
          mbar_wait(bar_bempty(s), ((it / SB) & 1u) ^ 1u);
          mbar_arrive_expect_tx(bar_bfull(s), B_STAGE_BYTES);
          bulk_load_1d(b_ring + s * B_STAGE_BYTES, p.bimg + (size_t)st * KS * NC, B_STAGE_BYTES, bar_bfull(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    DENSE_REG_DEC();
    // D = f32, A = B = bf16, K-major both, M = 128; N = 64 (Xh x [Bh|Bl]) or 32 (Xl x Bh)
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    const uint32_t idesc64 = idesc_base | ((uint32_t)(NC >> 3) << 17);
    const uint32_t idesc32 = idesc_base | ((uint32_t)(NL >> 3) << 17);
    // B descriptor: K-major, no swizzle; LBO = NC * 16 B (next 8-wide K chunk), SBO = 128 B (next 8 columns), version 1
    const uint32_t desc_lo_const = (uint32_t)((NC * 16) >> 4) << 16;
    const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);
    uint32_t it = 0, item_idx = 0;
    const bool leader = elect_one();      // the whole role on one thread
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x, ++item_idx) {
      uint32_t rg, ks, st0, st1;
      item_range(item, rg, ks, st0, st1);
      mbar_wait(bar_accempty, (item_idx & 1u) ^ 1u);
      tc_fence_after();
      uint32_t acc_flag = 0;
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const int s = it % SB;
        mbar_wait(bar_bfull(s), (it / SB) & 1u);
        const uint32_t bsm = b_ring + s * B_STAGE_BYTES;
        const int slot = it % SLOTS;
        mbar_wait(bar_tfull(slot), (it / SLOTS) & 1u);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < RT; ++t) {
          const uint32_t d_t = tmem_base + D_COL0 + t * NC;
          const uint32_t a_t = tmem_base + A_COL0 + (slot * RT + t) * 32;
#pragma unroll
          for (int i = 0; i < KS / 16; ++i) {
            const uint32_t baddr = bsm + (uint32_t)(i * 16 * NC * 2);
            const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo_const | ((baddr >> 4) & 0x3FFFu));
            tc_mma_ts(d_t, a_t + 8 * i, bdesc, idesc64, acc_flag | (uint32_t)i);          // Xh x [Bh | Bl]
            tc_mma_ts(d_t, a_t + 16 + 8 * i, bdesc, idesc32, 1u);                          // Xl x Bh
          }
        }
        tc_commit(bar_tempty(slot));
        tc_commit(bar_bempty(s));
        acc_flag = 1;
      }
      tc_commit(bar_accfull);
    }
    __syncwarp();
  } else if (warp == 3) {
    DENSE_REG_DEC();
  } else {
    DENSE_REG_INC();
    const int tile = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t sw = (uint32_t)(row_in_tile & 7);      // SWIZZLE_128B (ROWS mode): 16-byte chunk c of row r at c ^ (r & 7)
    uint32_t it = 0, item_idx = 0;
    for (uint32_t item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_idx) {
      uint32_t rg, ks, st0, st1;
      item_range(item, rg, ks, st0, st1);
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const int s = it % SA;
        mbar_wait(bar_afull(s), (it / SA) & 1u);
        const uint32_t tbase = a_ring + s * A_STAGE_BYTES + tile * A_TILE_BYTES;
        float x[KS];
        if (COLS) {
          // tile [32 k][128 rows] floats: a warp reads 32 consecutive floats per k (conflict-free)
#pragma unroll
          for (int k = 0; k < KS; ++k)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[k]) : "r"(tbase + (uint32_t)(k * 512 + row_in_tile * 4)));
        } else {
          const uint32_t arow = tbase + row_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(x[4 * q]), "=f"(x[4 * q + 1]), "=f"(x[4 * q + 2]), "=f"(x[4 * q + 3])
                         : "r"(arow + (((uint32_t)q ^ sw) << 4)));
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
        if (lane == 0) mbar_arrive(bar_aempty(s));      // (every value of the stage is in registers: split2 consumed them)
        const int slot = it % SLOTS;
        mbar_wait(bar_tempty(slot), ((it / SLOTS) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_addr + A_COL0 + (slot * RT + tile) * 32;
        tmem_st16(ta, hi);
        tmem_st16(ta + 16, lo);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfull(slot));
      }
      // ---- epilogue
      const uint64_t r = (uint64_t)rg * (RT * 128) + tile * 128 + row_in_tile;
      const bool live = r < p.rows;
      float ar = 1.0f, br = 1.0f;
      if (!p.partial && live) {
        if (p.a) ar = __ldg(p.a + r);
        if (p.b) br = __ldg(p.b + r);
      }
      mbar_wait(bar_accfull, item_idx & 1u);
      tc_fence_after();
      uint32_t dh[32], dl[32];
      tmem_ld32(tmem_base + lane_addr + D_COL0 + tile * NC, dh);
      tmem_ld32(tmem_base + lane_addr + D_COL0 + tile * NC + 32, dl);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty);
      if (live) {
        if (p.partial) {
          float* dst = p.partial + ((uint64_t)ks * p.rows + r) * NL;
#pragma unroll
          for (int c = 0; c < NL; c += 4)
            *reinterpret_cast<float4*>(dst + c) =
                make_float4(__uint_as_float(dh[c]) + __uint_as_float(dl[c]), __uint_as_float(dh[c + 1]) + __uint_as_float(dl[c + 1]),
                            __uint_as_float(dh[c + 2]) + __uint_as_float(dl[c + 2]), __uint_as_float(dh[c + 3]) + __uint_as_float(dl[c + 3]));
        } else {
          float* dst = p.out + r * p.ldo;
#pragma unroll
          for (int c = 0; c < NL; ++c)
            if ((uint32_t)c < p.l) dst[c] = ar * (__uint_as_float(dh[c]) + __uint_as_float(dl[c])) - br * cv_s[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// column sums cpart[block][c] = sum_k e_k Bin[k][c] in f64 (fixed order), 8 rows x 32 columns per CTA step
__global__ void __launch_bounds__(256) dense_colsum_kernel(const float* __restrict__ bin, uint64_t K, uint32_t l, uint32_t ld,
                                                           const float* __restrict__ e, double* __restrict__ cpart) {
  __shared__ double red[256];
  const int cidx = threadIdx.x & 31, rr = threadIdx.x >> 5;
  double acc = 0.0;
  if ((uint32_t)cidx < l)
    for (uint64_t k = (uint64_t)blockIdx.x * 8 + rr; k < K; k += (uint64_t)gridDim.x * 8)
      acc += (double)bin[k * ld + cidx] * (double)(e ? e[k] : 1.0f);
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rr == 0) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q * 32 + cidx];
    cpart[(uint64_t)blockIdx.x * 32 + cidx] = s;
  }
}
__global__ void dense_cvec_kernel(const double* __restrict__ cpart, int nparts, float* __restrict__ cvec) {
  const int cidx = threadIdx.x;
  if (cidx >= 32) return;
  double s = 0.0;
  for (int q = 0; q < nparts; ++q) s += cpart[(uint64_t)q * 32 + cidx];
  cvec[cidx] = (float)s;
}

// [Bh | Bl] image: MMA group g covers k = 16 g .. 16 g + 15 in natural order; element (slot s, column n in 0..63;
// n < 32: hi part of logical column n, n >= 32: lo part of column n - 32) at bf16 offset
//   g*16*NC + (s/8)*(8*NC) + (n/8)*64 + (n%8)*8 + (s%8)            (UMMA K-major core matrices, no swizzle)
__global__ void __launch_bounds__(256) dense_prep_b_kernel(const float* __restrict__ bin, uint64_t K, uint64_t Kpad,
                                                           uint32_t l, uint32_t ld, const float* __restrict__ f,
                                                           __nv_bfloat16* __restrict__ img) {
  const uint64_t total = (Kpad / 8) * NL;       // one thread per (8-slot K chunk, logical column)
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t n = (uint32_t)(t % NL);
    const uint64_t kc = t / NL;
    const uint64_t g = kc >> 1;
    const uint32_t half_idx = (uint32_t)(kc & 1);
    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint64_t k = g * 16 + half_idx * 8 + kk;
      float v = 0.0f;
      if (k < K && n < l) {
        v = bin[k * ld + n];
        if (f) v *= f[k];
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      vh[kk] = h;
      vl[kk] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    const uint64_t off = g * 16 * NC + (uint64_t)half_idx * (8 * NC) + (n & 7) * 8;
    *reinterpret_cast<uint4*>(img + off + (n >> 3) * 64) = *reinterpret_cast<const uint4*>(vh);
    *reinterpret_cast<uint4*>(img + off + ((n + NL) >> 3) * 64) = *reinterpret_cast<const uint4*>(vl);
  }
}

__global__ void __launch_bounds__(256) dense_reduce_kernel(const float* __restrict__ partial, int nsplit, uint64_t rows,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ cvec, float* __restrict__ out,
                                                           uint32_t ldo, uint32_t l) {
  const uint64_t total = rows * NL;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / NL;
    const uint32_t cc = (uint32_t)(t % NL);
    if (cc >= l) continue;
    float s = 0.0f;
    for (int q = 0; q < nsplit; ++q) s += partial[((uint64_t)q * rows + r) * NL + cc];
    out[r * ldo + cc] = (a ? a[r] : 1.0f) * s - (b ? b[r] : 1.0f) * cvec[cc];
  }
}

// plain fp32 fallback for shapes below one tile (tests with a handful of condensed features) and for l > 32
__global__ void __launch_bounds__(256) dense_simt_kernel(const float* __restrict__ cmat, uint32_t ldc, bool cols,
                                                         uint64_t rows, uint64_t K, const float* __restrict__ bin,
                                                         uint32_t l, uint32_t ld, const float* __restrict__ f,
                                                         const float* __restrict__ e, const float* __restrict__ a,
                                                         const float* __restrict__ b, float* __restrict__ out, uint32_t ldo) {
  const uint64_t total = rows * l;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / l;
    const uint32_t cc = (uint32_t)(t - r * l);
    double acc = 0.0, cv = 0.0;
    for (uint64_t k = 0; k < K; ++k) {
      const float x = cols ? cmat[k * ldc + r] : cmat[r * ldc + k];
      const float bv = bin[k * ld + cc];
      acc += (double)x * (double)(bv * (f ? f[k] : 1.0f));
      cv += (double)bv * (double)(e ? e[k] : 1.0f);
    }
    out[r * ldo + cc] = (float)((double)(a ? a[r] : 1.0f) * acc - (double)(b ? b[r] : 1.0f) * cv);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_dense() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
}  // namespace

int launch_dense_product(gpca_ctx* c, const DenseProduct& p) {
  const uint64_t rows = p.cols_mode ? p.R : p.N, K = p.cols_mode ? p.N : p.R;
  if (rows == 0 || K == 0 || p.l == 0) return GPCA_OK;
  EncodeTiledFn enc = get_encode_fn_dense();
  const bool tensor = enc && p.l <= NL && rows >= 128 && K >= KS && (p.ldc % 4) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && rows < (1ull << 31) && K < (1ull << 31) &&
                      !getenv("GPCA_DEBUG_DENSE_SIMT");
  if (!tensor) {
    const uint64_t total = rows * p.l;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 16);
    dense_simt_kernel<<<grid, 256, 0, c->stream>>>(p.C, p.ldc, p.cols_mode, rows, K, p.Bin, p.l, p.ld, p.f, p.e, p.a, p.b,
                                                   p.out, p.ldo);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
    return GPCA_OK;
  }
  const uint64_t Kpad = round_up(K, KS);
  const uint32_t total_stages = (uint32_t)(Kpad / KS);
  GPCA_CUDA_TRY(c, c->ws_bytes.alloc(Kpad * NC * 2));
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(c->ws_bytes.p);
  int nb = (int)std::min<uint64_t>((K + 63) / 64, (uint64_t)c->sm_count * 4);
  if (nb < 1) nb = 1;
  GPCA_CUDA_TRY(c, c->ws_cpart.alloc((size_t)nb * 32));
  GPCA_CUDA_TRY(c, c->ws_cvec.alloc(64 + 8));
  float* cvec = c->ws_cvec.p;
  dense_colsum_kernel<<<nb, 256, 0, c->stream>>>(p.Bin, K, p.l, p.ld, p.e, c->ws_cpart.p);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  dense_cvec_kernel<<<1, 32, 0, c->stream>>>(c->ws_cpart.p, nb, cvec);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  {
    const uint64_t total = (Kpad / 8) * NL;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 8);
    dense_prep_b_kernel<<<grid, 256, 0, c->stream>>>(p.Bin, K, Kpad, p.l, p.ld, p.f, img);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  const uint32_t row_groups = (uint32_t)((rows + RT * 128 - 1) / (RT * 128));
  const uint32_t slots = (uint32_t)c->sm_count * 2;
  // K split: enough items for about four rounds over the persistent CTAs, at least 16 stages (512 k) per item
  uint32_t ksplit = 1;
  if (row_groups < 4 * slots) {
    ksplit = (4 * slots + row_groups - 1) / row_groups;
    const uint32_t max_split = std::max<uint32_t>(1u, total_stages / 16);
    ksplit = std::min(ksplit, max_split);
  }
  const uint32_t spp = (total_stages + ksplit - 1) / ksplit;
  ksplit = (total_stages + spp - 1) / spp;
  DenseParams tp;
  tp.bimg = img;
  tp.rows = rows;
  tp.total_stages = total_stages;
  tp.stages_per_split = spp;
  tp.ksplit = ksplit;
  tp.row_groups = row_groups;
  tp.n_items = row_groups * ksplit;
  tp.a = p.a;
  tp.b = p.b;
  tp.cvec = cvec;
  tp.out = p.out;
  tp.ldo = p.ldo;
  tp.l = p.l;
  tp.partial = nullptr;
  if (ksplit > 1) {
    GPCA_CUDA_TRY(c, c->ws_partial.alloc((size_t)ksplit * rows * NL));
    tp.partial = c->ws_partial.p;
  }
  CUtensorMap tmap;
  {
    // C [N x R] fp32, row stride ldc: dimension 0 = condensed feature (contiguous), dimension 1 = sample
    const cuuint64_t dims[2] = {(cuuint64_t)p.R, (cuuint64_t)p.N};
    const cuuint64_t strides[1] = {(cuuint64_t)p.ldc * 4};
    const cuuint32_t box_rows[2] = {KS, 128};        // ROWS: 32 k x 128 samples, 128-byte swizzle
    const cuuint32_t box_cols[2] = {128, KS};        // COLS: 128 features x 32 k (samples), no swizzle
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p.C, dims, strides,
                           p.cols_mode ? box_cols : box_rows, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           p.cols_mode ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      c->set_error("dense_tc: cuTensorMapEncodeTiled failed");
      return GPCA_ERR_CUDA;
    }
  }
  const uint32_t grid = tp.n_items < slots ? tp.n_items : slots;
  if (p.cols_mode) {
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(dense_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dense_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(tmap, tp);
  } else {
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(dense_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    dense_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(tmap, tp);
  }
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  if (ksplit > 1) {
    const uint64_t total = rows * NL;
    const int g2 = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 8);
    dense_reduce_kernel<<<g2, 256, 0, c->stream>>>(tp.partial, (int)ksplit, rows, p.a, p.b, cvec, p.out, p.ldo, p.l);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  return GPCA_OK;
}
