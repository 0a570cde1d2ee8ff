// kernels_dense.cu -- the small dense linear algebra around the sketch passes:
// Gaussian test matrices, l x l Gram (f64 accumulation), right-multiplication by an l x l2
// transform, single-CTA Jacobi eigensolver.  These replace the QR / small-SVD calls that
// efficient_pca makes through faer/LAPACK (external crate; call sites src/main.rs:648-659, :365).
#include "kernels.cuh"
#include "philox.cuh"

#define KLAUNCH_CHECK(c)                    \
  do {                                      \
    (c)->launches++;                        \
    GPCA_CUDA_TRY((c), cudaGetLastError()); \
  } while (0)

// ------------------------------------------------------------------------------------------
__global__ void gaussian_kernel(float* __restrict__ out, uint64_t rows, uint32_t cols, uint32_t ld, uint64_t seed,
                                uint32_t stream, uint64_t row0) {
  const uint32_t groups = (ld + 3) / 4;          // one thread per (row, group of 4 columns): one Philox call
  const uint64_t total = rows * groups;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / groups;
    const uint32_t cg = (uint32_t)(t - r * groups);
    float z[4];
    philox_normal4(seed, stream, row0 + r, cg, z);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t cidx = cg * 4 + j;
      if (cidx < ld) out[r * ld + cidx] = (cidx < cols) ? z[j] : 0.0f;
    }
  }
}

int launch_gaussian(gpca_ctx* c, float* d_out, uint64_t rows, uint32_t cols, uint32_t ld, uint64_t seed,
                    uint32_t stream, uint64_t row0) {
  const uint64_t total = rows * ((ld + 3) / 4);
  if (total == 0) return GPCA_OK;
  const int threads = 256;
  const uint64_t blocks = (total + threads - 1) / threads;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  gaussian_kernel<<<grid, threads, 0, c->stream>>>(d_out, rows, cols, ld, seed, stream, row0);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int stats_buffer(gpca_ctx* c, double** cpart, unsigned int** amax) {
  const size_t need = (size_t)STATS_MAX_PARTS * 32 + 2;
  if (c->ws_stats.n < need) {
    GPCA_CUDA_TRY(c, c->ws_stats.alloc(need));
    GPCA_CUDA_TRY(c, cudaMemsetAsync(c->ws_stats.p, 0, need * sizeof(double), c->stream));
    c->stats_pending = false;
  }
  *cpart = c->ws_stats.p;
  *amax = reinterpret_cast<unsigned int*>(c->ws_stats.p + (size_t)STATS_MAX_PARTS * 32);
  return GPCA_OK;
}

// call before a producer accumulates into the max-abs word: statistics that nobody consumed (other engine, error
// path) must not leak into the new ones
int stats_begin_produce(gpca_ctx* c, unsigned int* amax) {
  if (c->stats_pending) GPCA_CUDA_TRY(c, cudaMemsetAsync(amax, 0, sizeof(unsigned int), c->stream));
  c->stats_pending = true;
  c->stats_for = nullptr;
  return GPCA_OK;
}

// Gaussian matrix + max |f o out| as a by-product (what the quantisation of the pass that consumes it needs).
// Same element -> (row, column group) mapping as gaussian_kernel.
__global__ void __launch_bounds__(256) gaussian_stats_kernel(float* __restrict__ out, uint64_t rows, uint32_t cols,
                                                             uint32_t ld, uint64_t seed, uint32_t stream, uint64_t row0,
                                                             const float* __restrict__ f,
                                                             unsigned int* __restrict__ amax_bits) {
  const uint32_t groups = (ld + 3) / 4;
  const uint64_t total = rows * groups;
  const bool vec4 = (ld & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  float mx = 0.0f;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / groups;
    const uint32_t cg = (uint32_t)(t - r * groups);
    float z[4];
    philox_normal4(seed, stream, row0 + r, cg, z);
    const float fr = f ? f[r] : 1.0f;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t cidx = cg * 4 + j;
      v[j] = (cidx < cols) ? z[j] : 0.0f;
      mx = fmaxf(mx, fabsf(v[j] * fr));
    }
    if (vec4) {      // rows padded to a multiple of 4 floats: one 16-byte store per thread
      *reinterpret_cast<float4*>(out + r * ld + cg * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (cg * 4 + j < ld) out[r * ld + cg * 4 + j] = v[j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0f && isfinite(mx)) atomicMax(amax_bits, __float_as_uint(mx));
}

int launch_gaussian_with_stats(gpca_ctx* c, float* d_out, uint64_t rows, uint32_t cols, uint32_t ld, uint64_t seed,
                               uint32_t stream, uint64_t row0, const float* d_f, const float* d_e) {
  (void)d_e;   // (the column sums are computed where the operand is quantised)
  c->stats_for = nullptr;
  if (cols > 32 || ld > 32 || rows == 0 || getenv("GPCA_DEBUG_NO_GAUSS_STATS"))
    return launch_gaussian(c, d_out, rows, cols, ld, seed, stream, row0);
  double* cpart = nullptr;
  unsigned int* amax = nullptr;
  GPCA_TRY(stats_buffer(c, &cpart, &amax));
  GPCA_TRY(stats_begin_produce(c, amax));
  const uint64_t total = rows * ((ld + 3) / 4);
  const uint64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  gaussian_stats_kernel<<<grid, 256, 0, c->stream>>>(d_out, rows, cols, ld, seed, stream, row0, d_f, amax);
  KLAUNCH_CHECK(c);
  c->stats_for = d_out;
  c->stats_l = cols;
  c->stats_nparts = 1;
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Gram: each CTA owns a contiguous range of rows, stages 64 rows at a time in shared memory and accumulates an
// LP x LP (LP = 32 or 64, padded) f64 partial with a 4 x 4 register tile per thread.  For LP = 32 the 256 threads form
// 4 groups that split the rows of a tile (all threads busy); the groups are summed through shared memory.  Partials are
// then summed in a fixed order by gram_reduce_kernel (deterministic).
constexpr int GRAM_ROWS = 64;

template <int LP>
__device__ __forceinline__ void gram_partial_body(const float* __restrict__ y, uint32_t l, uint32_t ld,
                                                  const uint64_t r_begin, const uint64_t r_end,
                                                  double* __restrict__ p) {
  // 4 x 4 output tiles, one per thread; only the ceil(l/4)^2 tiles that hold live columns are handed out, and the
  // warps left over split the rows of a tile between them (8 row groups for l <= 20, 4 for l <= 32, ...) -- with the
  // fixed 8 x 8 layout a 17-column Gram (EigenSNP's local bases) did 2.5x the useful FMAs.
  constexpr int TROWS = (LP == 32) ? GRAM_ROWS : GRAM_ROWS / 2;      // rows staged at a time (f64: 17 KB either way)
  // the tile is widened to f64 once, while it is staged (the fp32 -> f64 conversion is a quarter-rate instruction: with
  // fp32 tiles every thread converted 8 values per 16 FMAs and the kernel was conversion-bound)
  __shared__ __align__(16) double tile[TROWS][LP + 2];
  __shared__ double gsum[3072];                 // (groups - 1) x tiles x 16 <= 3072 for every split below
  const int tr = ((int)l + 3) / 4;              // live tiles per dimension
  const int tiles = tr * tr;
  const int wpg = (tiles + 31) / 32;            // warps per row group: 1, 2, 3, 4, ... 8
  const int groups = 8 / wpg;                   // 8, 4, 2, 2, 1 ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / wpg;
  const int t = (warp % wpg) * 32 + lane;
  const bool active = grp < groups && t < tiles;
  const int ti = active ? t / tr : 0, tj = active ? t % tr : 0;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // The next tile's global loads are in flight while the current one is multiplied (registers -> f64 tile after the
  // barrier): with load -> barrier -> multiply -> barrier in sequence the kernel ran at the global-memory latency
  // (0.4 TB/s on the 37,500 x 17 per-block matrices of EigenSNP's local bases).
  constexpr int NLD = TROWS * LP / 256;
  float stg[NLD];
  auto fetch = [&](uint64_t r0) {
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int e = threadIdx.x + q * 256;
      const int rr = e / LP, cc = e % LP;
      const uint64_t r = r0 + rr;
      stg[q] = (r < r_end && (uint32_t)cc < l) ? __ldg(y + r * ld + cc) : 0.0f;
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (uint64_t r0 = r_begin; r0 < r_end; r0 += TROWS) {
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int e = threadIdx.x + q * 256;
      tile[e / LP][e % LP] = (double)stg[q];
    }
    __syncthreads();
    if (r0 + TROWS < r_end) fetch(r0 + TROWS);
    if (active) {
#pragma unroll 4
      for (int rr = grp; rr < TROWS; rr += groups) {
        const double2 a0 = *reinterpret_cast<const double2*>(&tile[rr][ti * 4]);
        const double2 a1 = *reinterpret_cast<const double2*>(&tile[rr][ti * 4 + 2]);
        const double2 b0 = *reinterpret_cast<const double2*>(&tile[rr][tj * 4]);
        const double2 b1 = *reinterpret_cast<const double2*>(&tile[rr][tj * 4 + 2]);
        const double av[4] = {a0.x, a0.y, a1.x, a1.y};
        const double bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  if (groups > 1) {
    if (active && grp > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gsum[((grp - 1) * tiles + t) * 16 + i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (active && grp == 0) {
      for (int g = 0; g < groups - 1; ++g)      // fixed order: deterministic
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += gsum[(g * tiles + t) * 16 + i * 4 + j];
    }
  }
  if (active && grp == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) p[(ti * 4 + i) * LP + tj * 4 + j] = acc[i][j];
  }
}

template <int LP>
__global__ void __launch_bounds__(256, 3) gram_partial_kernel(const float* __restrict__ y, uint64_t n, uint32_t l,
                                                           uint32_t ld, uint64_t rows_per_cta,
                                                           double* __restrict__ partial) {
  const uint64_t r_begin = blockIdx.x * rows_per_cta;
  uint64_t r_end = r_begin + rows_per_cta;
  if (r_end > n) r_end = n;
  gram_partial_body<LP>(y, l, ld, r_begin, r_end, partial + (uint64_t)blockIdx.x * LP * LP);
}

// batched: grid = (parts, problems); partial of (problem b, part q) at [(b * parts + q) * 1024]
__global__ void __launch_bounds__(256, 3) gram_batch_kernel(const float* __restrict__ base, uint32_t ld,
                                                         const DenseProb* __restrict__ probs,
                                                         double* __restrict__ partial) {
  const DenseProb pb = probs[blockIdx.y];
  uint64_t per = (pb.rows + gridDim.x - 1) / gridDim.x;
  per = (per + GRAM_ROWS - 1) / GRAM_ROWS * GRAM_ROWS;
  uint64_t r_begin = (uint64_t)blockIdx.x * per;
  uint64_t r_end = r_begin + per;
  if (r_begin > pb.rows) r_begin = pb.rows;
  if (r_end > pb.rows) r_end = pb.rows;
  gram_partial_body<32>(base + pb.off, pb.l, ld, r_begin, r_end,
                        partial + ((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * 1024);
}

__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ partial, int nparts, uint32_t l,
                                                          int lp, double* __restrict__ g) {
  // one warp per output element: lane q sums the partials q, q+32, ... in index order, then a fixed butterfly
  // (deterministic; a single thread walking hundreds of partials 8 KB apart took 80 us)
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int t = w; t < (int)(l * l); t += nw) {
    const int i = t / l, j = t % l;
    double s = 0.0;
    for (int p = lane; p < nparts; p += 32) s += partial[(uint64_t)p * lp * lp + i * lp + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) g[t] = s;
  }
}

// Cross Gram  G = A^T B  (A, B [n x l] fp32 with the same row stride), f64 accumulation; same tiling as the Gram.
template <int LP>
__global__ void __launch_bounds__(256) cross_gram_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                 uint64_t n, uint32_t l, uint32_t ld,
                                                                 uint64_t rows_per_cta, double* __restrict__ partial) {
  constexpr int TPD = LP / 4;
  constexpr int GROUPS = 256 / (TPD * TPD);
  __shared__ __align__(16) double ta[GRAM_ROWS / 2][LP + 2];      // (32-row tiles: two f64 tiles + the group sums in 48 KB)
  __shared__ __align__(16) double tb[GRAM_ROWS / 2][LP + 2];
  __shared__ double gsum[(GROUPS > 1) ? (GROUPS - 1) * LP * LP : 1];
  const int grp = threadIdx.x / (TPD * TPD);
  const int t = threadIdx.x % (TPD * TPD);
  const int ti = t / TPD, tj = t % TPD;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const uint64_t r_begin = blockIdx.x * rows_per_cta;
  uint64_t r_end = r_begin + rows_per_cta;
  if (r_end > n) r_end = n;
  for (uint64_t r0 = r_begin; r0 < r_end; r0 += GRAM_ROWS / 2) {
    {
      constexpr int NLD = (GRAM_ROWS / 2) * LP / 256;
      float sa[NLD], sb[NLD];
#pragma unroll
      for (int q = 0; q < NLD; ++q) {
        const int e = threadIdx.x + q * 256;
        const int rr = e / LP, cc = e % LP;
        const uint64_t r = r0 + rr;
        const bool live = r < r_end && (uint32_t)cc < l;
        sa[q] = live ? __ldg(a + r * ld + cc) : 0.0f;
        sb[q] = live ? __ldg(b + r * ld + cc) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < NLD; ++q) {
        const int e = threadIdx.x + q * 256;
        ta[e / LP][e % LP] = (double)sa[q];
        tb[e / LP][e % LP] = (double)sb[q];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = grp; rr < GRAM_ROWS / 2; rr += GROUPS) {
      const double2 a0 = *reinterpret_cast<const double2*>(&ta[rr][ti * 4]);
      const double2 a1 = *reinterpret_cast<const double2*>(&ta[rr][ti * 4 + 2]);
      const double2 b0 = *reinterpret_cast<const double2*>(&tb[rr][tj * 4]);
      const double2 b1 = *reinterpret_cast<const double2*>(&tb[rr][tj * 4 + 2]);
      const double av[4] = {a0.x, a0.y, a1.x, a1.y};
      const double bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (GROUPS > 1) {
    if (grp > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gsum[((grp - 1) * LP + ti * 4 + i) * LP + tj * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          for (int g = 0; g < GROUPS - 1; ++g) acc[i][j] += gsum[(g * LP + ti * 4 + i) * LP + tj * 4 + j];
    }
  }
  if (grp == 0) {
    double* p = partial + (uint64_t)blockIdx.x * LP * LP;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) p[(ti * 4 + i) * LP + tj * 4 + j] = acc[i][j];
  }
}

int launch_cross_gram(gpca_ctx* c, const float* d_a, const float* d_b, uint64_t n, uint32_t l, uint32_t ld,
                      double* d_g) {
  if (l == 0 || l > 64) {
    c->set_error("launch_cross_gram: l must be in 1..64");
    return GPCA_ERR_INVALID;
  }
  const int lp = (l <= 32) ? 32 : 64;
  int nparts = (int)((n + 255) / 256);
  if (nparts > c->sm_count * 4) nparts = c->sm_count * 4;
  if (nparts < 1) nparts = 1;
  uint64_t rows_per_cta = (n + nparts - 1) / nparts;
  rows_per_cta = round_up(rows_per_cta ? rows_per_cta : 1, GRAM_ROWS);
  nparts = (int)((n + rows_per_cta - 1) / rows_per_cta);
  if (nparts < 1) nparts = 1;
  GPCA_CUDA_TRY(c, c->ws_gram.alloc((size_t)nparts * lp * lp + 4096));
  if (lp == 32)
    cross_gram_partial_kernel<32><<<nparts, 256, 0, c->stream>>>(d_a, d_b, n, l, ld, rows_per_cta, c->ws_gram.p);
  else
    cross_gram_partial_kernel<64><<<nparts, 256, 0, c->stream>>>(d_a, d_b, n, l, ld, rows_per_cta, c->ws_gram.p);
  KLAUNCH_CHECK(c);
  gram_reduce_kernel<<<(l * l + 7) / 8, 256, 0, c->stream>>>(c->ws_gram.p, nparts, l, lp, d_g);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_gram(gpca_ctx* c, const float* d_y, uint64_t n, uint32_t l, uint32_t ld, double* d_g) {
  if (l == 0 || l > 64) {
    c->set_error("launch_gram: l must be in 1..64");
    return GPCA_ERR_INVALID;
  }
  const int lp = (l <= 32) ? 32 : 64;
  int nparts = (int)((n + 255) / 256);
  if (nparts > c->sm_count * 4) nparts = c->sm_count * 4;
  if (nparts < 1) nparts = 1;
  uint64_t rows_per_cta = (n + nparts - 1) / nparts;
  rows_per_cta = round_up(rows_per_cta ? rows_per_cta : 1, GRAM_ROWS);
  nparts = (int)((n + rows_per_cta - 1) / rows_per_cta);
  if (nparts < 1) nparts = 1;
  GPCA_CUDA_TRY(c, c->ws_gram.alloc((size_t)nparts * lp * lp + 4096));
  if (lp == 32)
    gram_partial_kernel<32><<<nparts, 256, 0, c->stream>>>(d_y, n, l, ld, rows_per_cta, c->ws_gram.p);
  else
    gram_partial_kernel<64><<<nparts, 256, 0, c->stream>>>(d_y, n, l, ld, rows_per_cta, c->ws_gram.p);
  KLAUNCH_CHECK(c);
  gram_reduce_kernel<<<(l * l + 7) / 8, 256, 0, c->stream>>>(c->ws_gram.p, nparts, l, lp, d_g);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// out[r, :l2] = y[r, :l] * T   (T f64 in shared memory, f64 accumulation, fp32 store).
// 128-row tiles, widened to f64 while they are staged.  Thread (row pair rr = tid/4 and rr + 64, column lane cg = tid%4)
// owns the column pairs 2cg + 8j + {0,1}, j < NOWN/2, of both rows: one 16-byte shared-memory load of T feeds four
// FMAs (the first version -- one row, one 8-byte load per FMA, an fp32 -> f64 conversion per element and thread -- was
// bound by the shared-memory pipe at 0.8 TB/s).
constexpr int AR_ROWS = 128;
template <int NOWN>
__global__ void __launch_bounds__(256) apply_right_kernel(const float* __restrict__ y, uint64_t n, uint32_t l,
                                                          uint32_t ld, const double* __restrict__ t, uint32_t l2,
                                                          float* __restrict__ out, uint32_t ldo) {
  static_assert(NOWN % 2 == 0, "columns are owned in pairs");
  extern __shared__ __align__(16) double sm[];
  const int l2p = NOWN * 4;                                    // padded output width
  double* ts = sm;                                             // [l][l2p] (zero padded)
  const int lp = (int)l + 1;
  double* tile = sm + l * l2p;                                 // [AR_ROWS][l+1]
  for (int i = threadIdx.x; i < (int)(l * l2p); i += 256) {
    const int cc = i / l2p, c2 = i % l2p;
    ts[i] = ((uint32_t)c2 < l2) ? t[cc * l2 + c2] : 0.0;
  }
  const uint64_t ntiles = (n + AR_ROWS - 1) / AR_ROWS;
  const bool vec2 = (ldo & 1u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
  // warp = row, lane = column; for l <= 32 the next tile's loads are in flight while the current tile is multiplied
  float stg[AR_ROWS / 8];
  auto fetch = [&](uint64_t r0, uint32_t c0) {
    const uint32_t cc = c0 + (threadIdx.x & 31);
#pragma unroll
    for (int q = 0; q < AR_ROWS / 8; ++q) {
      const uint64_t r = r0 + (threadIdx.x >> 5) + 8 * q;
      stg[q] = (r < n && cc < l) ? __ldg(y + r * ld + cc) : 0.0f;
    }
  };
  if (l <= 32 && blockIdx.x < ntiles) fetch((uint64_t)blockIdx.x * AR_ROWS, 0);
  for (uint64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const uint64_t r0 = tix * AR_ROWS;
    __syncthreads();
    if (l <= 32) {      // registers prefetched during the previous tile's multiply -> f64 tile
      const uint32_t cc = threadIdx.x & 31;
      if (cc < l) {
#pragma unroll
        for (int q = 0; q < AR_ROWS / 8; ++q) tile[((threadIdx.x >> 5) + 8 * q) * lp + cc] = (double)stg[q];
      }
    } else {
      for (uint32_t c0 = 0; c0 < l; c0 += 32) {
        fetch(r0, c0);
        const uint32_t cc = c0 + (threadIdx.x & 31);
        if (cc < l) {
#pragma unroll
          for (int q = 0; q < AR_ROWS / 8; ++q) tile[((threadIdx.x >> 5) + 8 * q) * lp + cc] = (double)stg[q];
        }
      }
    }
    __syncthreads();
    if (l <= 32 && tix + gridDim.x < ntiles) fetch((tix + gridDim.x) * AR_ROWS, 0);
    const int rr = threadIdx.x >> 2, cg = threadIdx.x & 3;
    double acc0[NOWN], acc1[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; ++j) acc0[j] = acc1[j] = 0.0;
    const double* y0 = tile + rr * lp;
    const double* y1 = tile + (rr + 64) * lp;
    for (uint32_t cc = 0; cc < l; ++cc) {
      const double a0 = y0[cc], a1 = y1[cc];
      const double* trow = ts + cc * l2p + 2 * cg;
#pragma unroll
      for (int j = 0; j < NOWN / 2; ++j) {
        const double2 tv = *reinterpret_cast<const double2*>(trow + 8 * j);
        acc0[2 * j] = fma(a0, tv.x, acc0[2 * j]);
        acc0[2 * j + 1] = fma(a0, tv.y, acc0[2 * j + 1]);
        acc1[2 * j] = fma(a1, tv.x, acc1[2 * j]);
        acc1[2 * j + 1] = fma(a1, tv.y, acc1[2 * j + 1]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t r = r0 + rr + 64 * h;
      if (r >= n) continue;
      const double* acc = h ? acc1 : acc0;
      float* orow = out + r * ldo;
#pragma unroll
      for (int j = 0; j < NOWN / 2; ++j) {
        const uint32_t c0 = 2 * cg + 8 * j;
        if (vec2 && c0 + 1 < l2) {
          *reinterpret_cast<float2*>(orow + c0) = make_float2((float)acc[2 * j], (float)acc[2 * j + 1]);
        } else {
          if (c0 < l2) orow[c0] = (float)acc[2 * j];
          if (c0 + 1 < l2) orow[c0 + 1] = (float)acc[2 * j + 1];
        }
      }
    }
  }
}

template <int NOWN>
static int run_apply_right(gpca_ctx* c, const float* d_y, uint64_t n, uint32_t l, uint32_t ld, const double* d_t,
                           uint32_t l2, float* d_out, uint32_t ldo) {
  const size_t smem = ((size_t)l * NOWN * 4 + (size_t)AR_ROWS * (l + 1)) * sizeof(double);
  if (smem > 48 * 1024)
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(apply_right_kernel<NOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint64_t ntiles = (n + AR_ROWS - 1) / AR_ROWS;
  const int grid = (int)(ntiles < (uint64_t)c->sm_count * 4 ? ntiles : (uint64_t)c->sm_count * 4);
  apply_right_kernel<NOWN><<<grid, 256, smem, c->stream>>>(d_y, n, l, ld, d_t, l2, d_out, ldo);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_apply_right(gpca_ctx* c, const float* d_y, uint64_t n, uint32_t l, uint32_t ld, const double* d_t,
                       uint32_t l2, float* d_out, uint32_t ldo) {
  if (n == 0 || l == 0 || l2 == 0) return GPCA_OK;
  const uint32_t need = (l2 + 3) / 4;
  if (need <= 2) return run_apply_right<2>(c, d_y, n, l, ld, d_t, l2, d_out, ldo);
  if (need <= 6) return run_apply_right<6>(c, d_y, n, l, ld, d_t, l2, d_out, ldo);
  if (need <= 8) return run_apply_right<8>(c, d_y, n, l, ld, d_t, l2, d_out, ldo);
  if (need <= 12) return run_apply_right<12>(c, d_y, n, l, ld, d_t, l2, d_out, ldo);
  return run_apply_right<16>(c, d_y, n, l, ld, d_t, l2, d_out, ldo);
}

// ------------------------------------------------------------------------------------------
// Single-CTA two-sided cyclic Jacobi (round-robin pairing), f64, l <= 64.
__device__ __forceinline__ void jacobi_eigh_body(const double* __restrict__ a_in, uint32_t l,
                                                 double* __restrict__ evals, double* __restrict__ evecs,
                                                 double* jsm) {
  constexpr int LP = 64;
  double (*A)[LP + 1] = reinterpret_cast<double (*)[LP + 1]>(jsm);
  double (*V)[LP + 1] = reinterpret_cast<double (*)[LP + 1]>(jsm + LP * (LP + 1));
  __shared__ double cs[LP / 2][2];
  __shared__ int pairs[LP / 2][2];
  __shared__ int order[LP];
  __shared__ double red[8];
  __shared__ int done;
  const int n = (int)l;
  const int ne = n + (n & 1);  // even size, dummy index n if odd
  for (int t = threadIdx.x; t < LP * LP; t += blockDim.x) {
    const int i = t / LP, j = t % LP;
    A[i][j] = (i < n && j < n) ? 0.5 * (a_in[i * n + j] + a_in[j * n + i]) : 0.0;
    V[i][j] = (i == j) ? 1.0 : 0.0;
  }
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  const int npairs = ne / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence: off-diagonal Frobenius^2 vs diagonal
    double off = 0.0, dg = 0.0;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      const double v = A[i][j];
      if (i == j) dg += v * v;
      else off += v * v;
    }
    off = warp_sum_f64(off);
    dg = warp_sum_f64(dg);
    if ((threadIdx.x & 31) == 0) {
      red[(threadIdx.x >> 5)] = off;
    }
    __syncthreads();
    double off_t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) off_t += red[w];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5)] = dg;
    __syncthreads();
    double dg_t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dg_t += red[w];
    __syncthreads();
    if (off_t <= 1e-28 * dg_t || off_t == 0.0) break;
    for (int round = 0; round < ne - 1; ++round) {
      // round-robin tournament: position 0 fixed, others rotate
      if ((int)threadIdx.x < npairs) {
        const int k = threadIdx.x;
        int p = (k == 0) ? 0 : 1 + (round + k - 1) % (ne - 1);
        int q = 1 + (round + ne - 1 - k - 1) % (ne - 1);
        if (p > q) {
          const int tmp = p;
          p = q;
          q = tmp;
        }
        pairs[k][0] = p;
        pairs[k][1] = q;
        double cth = 1.0, sth = 0.0;
        if (q < n) {
          const double apq = A[p][q];
          if (apq != 0.0) {
            const double app = A[p][p], aqq = A[q][q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            cth = 1.0 / sqrt(1.0 + tt * tt);
            sth = tt * cth;
          }
        }
        cs[k][0] = cth;
        cs[k][1] = sth;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int t = threadIdx.x; t < npairs * n; t += blockDim.x) {
        const int k = t / n, i = t % n;
        const int p = pairs[k][0], q = pairs[k][1];
        const double cth = cs[k][0], sth = cs[k][1];
        if (q < n && sth != 0.0) {
          const double aip = A[i][p], aiq = A[i][q];
          A[i][p] = cth * aip - sth * aiq;
          A[i][q] = sth * aip + cth * aiq;
          const double vip = V[i][p], viq = V[i][q];
          V[i][p] = cth * vip - sth * viq;
          V[i][q] = sth * vip + cth * viq;
        }
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int t = threadIdx.x; t < npairs * n; t += blockDim.x) {
        const int k = t / n, j = t % n;
        const int p = pairs[k][0], q = pairs[k][1];
        const double cth = cs[k][0], sth = cs[k][1];
        if (q < n && sth != 0.0) {
          const double apj = A[p][j], aqj = A[q][j];
          A[p][j] = cth * apj - sth * aqj;
          A[q][j] = sth * apj + cth * aqj;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // sort descending (rank by counting), write out
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double di = A[i][i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = A[j][j];
      if (dj > di || (dj == di && j < i)) ++rank;
    }
    order[rank] = i;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) evals[t] = A[order[t]][order[t]];
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t % n;
    evecs[i * n + j] = V[i][order[j]];
  }
}

__global__ void __launch_bounds__(256) jacobi_eigh_kernel(const double* __restrict__ a_in, uint32_t l,
                                                          double* __restrict__ evals, double* __restrict__ evecs,
                                                          const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag) return;   // the Cholesky path already produced the transform
  extern __shared__ double jsm[];
  jacobi_eigh_body(a_in, l, evals, evecs, jsm);
}

// batched: one CTA per problem; G_b at g + b*1024 (l x l compact), evals at + b*32, evecs at + b*1024
__global__ void __launch_bounds__(256) jacobi_eigh_batch_kernel(const double* __restrict__ g,
                                                                const DenseProb* __restrict__ probs,
                                                                double* __restrict__ evals, double* __restrict__ evecs,
                                                                const int* __restrict__ skip_flags) {
  const uint32_t b = blockIdx.x;
  if (skip_flags && skip_flags[b]) return;
  extern __shared__ double jsm[];
  jacobi_eigh_body(g + (size_t)b * 1024, probs[b].l, evals + (size_t)b * 32, evecs + (size_t)b * 1024, jsm);
}

int launch_jacobi_eigh(gpca_ctx* c, const double* d_a, uint32_t l, double* d_evals, double* d_evecs,
                       const int* d_skip_flag) {
  if (l == 0 || l > 64) {
    c->set_error("jacobi_eigh: l must be in 1..64");
    return GPCA_ERR_INVALID;
  }
  const size_t smem = 2 * 64 * 65 * sizeof(double);
  GPCA_CUDA_TRY(c, cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jacobi_eigh_kernel<<<1, 256, smem, c->stream>>>(d_a, l, d_evals, d_evecs, d_skip_flag);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void make_orth_transform_kernel(const double* __restrict__ evals, const double* __restrict__ evecs,
                                           uint32_t l, double* __restrict__ t, double rel_eps,
                                           const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag) return;
  const double lmax = evals[0];
  for (int i = threadIdx.x; i < (int)(l * l); i += blockDim.x) {
    const int j = i % l;
    const double lam = evals[j];
    t[i] = (lam > rel_eps * lmax && lam > 0.0) ? evecs[i] / sqrt(lam) : 0.0;
  }
}

int launch_make_orth_transform(gpca_ctx* c, const double* d_evals, const double* d_evecs, uint32_t l, double* d_t,
                               double rel_eps, const int* d_skip_flag) {
  make_orth_transform_kernel<<<1, 256, 0, c->stream>>>(d_evals, d_evecs, l, d_t, rel_eps, d_skip_flag);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Sign convention on the device: flags[j] = 1 when the largest-|.| entry of column j (first one on ties) is negative.
// Two steps: every CTA scans a strided set of rows with lane = column (coalesced row reads) and leaves its best
// (|value|, row) per column; one small CTA then picks the winner per column in CTA order.  Ties go to the smaller row
// index at every level, so the result does not depend on the grid.
struct SignCand {
  float a;
  uint32_t pad;
  unsigned long long i;
};
__device__ __forceinline__ bool sign_better(float a2, unsigned long long i2, float a1, unsigned long long i1) {
  return a2 > a1 || (a2 == a1 && i2 < i1);
}
__global__ void __launch_bounds__(256) sign_flags_partial_kernel(const float* __restrict__ x, uint64_t n, uint32_t k,
                                                                 uint32_t ld, SignCand* __restrict__ part) {
  __shared__ float sv[8][32];
  __shared__ unsigned long long si[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t gw = (uint64_t)blockIdx.x * 8 + w, nw = (uint64_t)gridDim.x * 8;
  for (uint32_t c0 = 0; c0 < k; c0 += 32) {
    const uint32_t j = c0 + lane;
    float best = -1.0f;
    unsigned long long bi = 0;
    if (j < k) {
      for (uint64_t i = gw; i < n; i += nw) {
        const float a = fabsf(x[i * ld + j]);
        if (a > best) {      // rows are visited in increasing order: the first maximum is kept
          best = a;
          bi = i;
        }
      }
    }
    sv[w][lane] = best;
    si[w][lane] = bi;
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int q = 1; q < 8; ++q)
        if (sign_better(sv[q][lane], si[q][lane], best, bi)) {
          best = sv[q][lane];
          bi = si[q][lane];
        }
      if (j < k) {
        SignCand cnd;
        cnd.a = best;
        cnd.pad = 0;
        cnd.i = bi;
        part[(uint64_t)blockIdx.x * 64 + j] = cnd;
      }
    }
    __syncthreads();
  }
}

__global__ void sign_flags_final_kernel(const float* __restrict__ x, uint64_t n, uint32_t k, uint32_t ld,
                                        const SignCand* __restrict__ part, int nparts, int* __restrict__ flags) {
  const uint32_t j = threadIdx.x;
  if (j >= k) return;
  float best = -1.0f;
  unsigned long long bi = 0;
  for (int q = 0; q < nparts; ++q) {
    const SignCand cnd = part[(uint64_t)q * 64 + j];
    if (sign_better(cnd.a, cnd.i, best, bi)) {
      best = cnd.a;
      bi = cnd.i;
    }
  }
  flags[j] = (n && x[bi * ld + j] < 0.0f) ? 1 : 0;
}

__global__ void apply_flags_kernel(const float* __restrict__ x, uint64_t n, uint32_t k, uint32_t ld,
                                   const int* __restrict__ flags, float* __restrict__ out_f32,
                                   double* __restrict__ out_f64) {
  const uint64_t total = n * k;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = t / k;
    const uint32_t j = (uint32_t)(t - i * k);
    float v = x[i * ld + j];
    if (flags[j]) v = -v;
    if (out_f32) out_f32[t] = v;
    if (out_f64) out_f64[t] = (double)v;
  }
}

int launch_sign_flags(gpca_ctx* c, const float* d_x, uint64_t n, uint32_t k, uint32_t ld, int* d_flags) {
  if (k == 0) return GPCA_OK;
  if (k > 64) {
    c->set_error("launch_sign_flags: k must be <= 64");
    return GPCA_ERR_INVALID;
  }
  int nparts = (int)std::min<uint64_t>((n + 63) / 64, (uint64_t)c->sm_count * 4);
  if (nparts < 1) nparts = 1;
  static_assert(sizeof(SignCand) == 2 * sizeof(double), "candidate records live in the f64 Gram workspace");
  GPCA_CUDA_TRY(c, c->ws_gram.alloc((size_t)nparts * 64 * 2));
  SignCand* part = reinterpret_cast<SignCand*>(c->ws_gram.p);
  sign_flags_partial_kernel<<<nparts, 256, 0, c->stream>>>(d_x, n, k, ld, part);
  KLAUNCH_CHECK(c);
  sign_flags_final_kernel<<<1, 64, 0, c->stream>>>(d_x, n, k, ld, part, nparts, d_flags);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_apply_flags(gpca_ctx* c, const float* d_x, uint64_t n, uint32_t k, uint32_t ld, const int* d_flags,
                       float* d_out_f32, double* d_out_f64) {
  const uint64_t total = n * k;
  if (total == 0) return GPCA_OK;
  const uint64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  apply_flags_kernel<<<grid, 256, 0, c->stream>>>(d_x, n, k, ld, d_flags, d_out_f32, d_out_f64);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}


// ------------------------------------------------------------------------------------------
// CholeskyQR transform: G = R^T R (upper R), T = R^-1, so that (Y T)^T (Y T) = I.  Single CTA, f64, l <= 64.
// ok_flag = 1 on success; 0 when a pivot is not safely positive (near rank deficiency) -- the caller then runs the
// eigen-based path (jacobi_eigh + make_orth_transform), which zeroes deficient directions instead.
// A (shared, symmetric, n x n valid) is overwritten by its upper Cholesky factor
__device__ __forceinline__ void chol_orth_body(double (*A)[65], int* bad_p, const int n, double* __restrict__ t,
                                               double rel_eps, int* __restrict__ ok_flag) {
  constexpr int LP = 64;
  int& bad = *bad_p;
  double dmax = 0.0;
  for (int i = 0; i < n; ++i) dmax = fmax(dmax, A[i][i]);
  const double thr = rel_eps * dmax;
  for (int k = 0; k < n; ++k) {
    const double piv = A[k][k];
    if (!(piv > thr) || !isfinite(piv)) {
      if (threadIdx.x == 0) bad = 1;
      break;                                     // uniform: every thread reads the same shared value
    }
    const double inv = 1.0 / sqrt(piv);
    __syncthreads();
    for (int j = k + (int)threadIdx.x; j < n; j += blockDim.x) A[k][j] *= inv;      // row k of R (incl. the diagonal)
    __syncthreads();
    const int m = n - k - 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j >= i) A[i][j] -= A[k][i] * A[k][j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) {
    if (threadIdx.x == 0) *ok_flag = 0;
    return;
  }
  // T = R^-1 (upper triangular), one thread per column by back substitution
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double col[LP];
    for (int i = n - 1; i >= 0; --i) {
      if (i > j) {
        col[i] = 0.0;
        continue;
      }
      double s = (i == j) ? 1.0 : 0.0;
      for (int mm = i + 1; mm <= j; ++mm) s -= A[i][mm] * col[mm];
      col[i] = s / A[i][i];
    }
    for (int i = 0; i < n; ++i) t[i * n + j] = col[i];
  }
  if (threadIdx.x == 0) *ok_flag = 1;
}

__global__ void __launch_bounds__(256) chol_orth_kernel(const double* __restrict__ g, uint32_t l, double* __restrict__ t,
                                                        double rel_eps, int* __restrict__ ok_flag) {
  constexpr int LP = 64;
  __shared__ double A[LP][LP + 1];   // 33 KB
  __shared__ int bad;
  const int n = (int)l;
  for (int e = threadIdx.x; e < LP * LP; e += blockDim.x) {
    const int i = e / LP, j = e % LP;
    A[i][j] = (i < n && j < n) ? 0.5 * (g[i * n + j] + g[j * n + i]) : 0.0;
  }
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  chol_orth_body(A, &bad, n, t, rel_eps, ok_flag);
}

// batched: one CTA per problem.  Sums the problem's Gram partials in a fixed order, stores the l x l Gram (for the
// eigen fallback) and the Cholesky transform T_b, flag_b.
__global__ void __launch_bounds__(256) chol_orth_batch_kernel(const double* __restrict__ partial, int nparts,
                                                              const DenseProb* __restrict__ probs,
                                                              double* __restrict__ g_out, double* __restrict__ t_out,
                                                              double rel_eps, int* __restrict__ ok_flags) {
  constexpr int LP = 64;
  __shared__ double A[LP][LP + 1];
  __shared__ int bad;
  const uint32_t b = blockIdx.x;
  const int n = (int)probs[b].l;
  for (int e = threadIdx.x; e < LP * LP; e += blockDim.x) {
    const int i = e / LP, j = e % LP;
    double s = 0.0;
    if (i < n && j < n) {
      for (int q = 0; q < nparts; ++q) {
        const double* pp = partial + ((uint64_t)b * nparts + q) * 1024;
        s += 0.5 * (pp[i * 32 + j] + pp[j * 32 + i]);
      }
      g_out[(size_t)b * 1024 + i * n + j] = s;
    }
    A[i][j] = s;
  }
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  chol_orth_body(A, &bad, n, t_out + (size_t)b * 1024, rel_eps, ok_flags + b);
}

int launch_chol_orth(gpca_ctx* c, const double* d_g, uint32_t l, double* d_t, double rel_eps, int* d_ok_flag) {
  if (l == 0 || l > 64) {
    c->set_error("chol_orth: l must be in 1..64");
    return GPCA_ERR_INVALID;
  }
  chol_orth_kernel<<<1, 256, 0, c->stream>>>(d_g, l, d_t, rel_eps, d_ok_flag);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Batched variants for the per-LD-block local bases (EigenSNP): one launch covers every block.
__global__ void gaussian_batch_kernel(float* __restrict__ base, uint32_t ld, const DenseProb* __restrict__ probs,
                                      uint64_t seed, const uint32_t* __restrict__ streams) {
  const DenseProb pb = probs[blockIdx.y];
  float* out = base + pb.off;
  const uint32_t stream = streams[blockIdx.y];
  const uint32_t groups = (ld + 3) / 4;
  const uint64_t total = (uint64_t)pb.rows * groups;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / groups;
    const uint32_t cg = (uint32_t)(t - r * groups);
    float z[4];
    philox_normal4(seed, stream, r, cg, z);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t cidx = cg * 4 + j;
      if (cidx < ld) out[r * ld + cidx] = (cidx < pb.l) ? z[j] : 0.0f;
    }
  }
}

int launch_gaussian_batch(gpca_ctx* c, float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                          uint64_t max_rows, uint64_t seed, const uint32_t* d_streams) {
  if (!n_probs || !max_rows) return GPCA_OK;
  const uint64_t total = max_rows * ((ld + 3) / 4);
  uint64_t gx = (total + 255) / 256;
  const uint64_t cap = ((uint64_t)c->sm_count * 32 + n_probs - 1) / n_probs;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, n_probs);
  gaussian_batch_kernel<<<grid, 256, 0, c->stream>>>(d_base, ld, d_probs, seed, d_streams);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int dense_batch_ws(gpca_ctx* c, uint32_t n_probs, DenseBatchWs& ws) {
  const size_t per = 3 * 1024 + 32 + 1;   // doubles per problem (flags stored in the last slot, as ints)
  GPCA_CUDA_TRY(c, c->ws_batch.alloc((size_t)n_probs * per));
  double* p = c->ws_batch.p;
  ws.G = p;
  ws.T = p + (size_t)n_probs * 1024;
  ws.evecs = p + (size_t)n_probs * 2048;
  ws.evals = p + (size_t)n_probs * 3072;
  ws.flags = reinterpret_cast<int*>(p + (size_t)n_probs * (3072 + 32));
  return GPCA_OK;
}

int launch_gram_batch(gpca_ctx* c, const float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                      uint64_t max_rows, int* nparts) {
  int np = (int)((max_rows + 2047) / 2048);
  const int cap = (c->sm_count * 16 + (int)n_probs - 1) / (int)n_probs;
  if (np > cap) np = cap;
  if (np > 64) np = 64;
  if (np < 1) np = 1;
  GPCA_CUDA_TRY(c, c->ws_gram.alloc((size_t)np * n_probs * 1024));
  dim3 grid((unsigned)np, n_probs);
  gram_batch_kernel<<<grid, 256, 0, c->stream>>>(d_base, ld, d_probs, c->ws_gram.p);
  KLAUNCH_CHECK(c);
  *nparts = np;
  return GPCA_OK;
}

int launch_chol_orth_batch(gpca_ctx* c, int nparts, const DenseProb* d_probs, uint32_t n_probs, double rel_eps,
                           const DenseBatchWs& ws) {
  chol_orth_batch_kernel<<<n_probs, 256, 0, c->stream>>>(c->ws_gram.p, nparts, d_probs, ws.G, ws.T, rel_eps, ws.flags);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_jacobi_eigh_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, const DenseBatchWs& ws,
                             bool use_flags) {
  const size_t smem = 2 * 64 * 65 * sizeof(double);
  GPCA_CUDA_TRY(c, cudaFuncSetAttribute(jacobi_eigh_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  jacobi_eigh_batch_kernel<<<n_probs, 256, smem, c->stream>>>(ws.G, d_probs, ws.evals, ws.evecs,
                                                              use_flags ? ws.flags : nullptr);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

__global__ void make_orth_transform_batch_kernel(const DenseProb* __restrict__ probs, const double* __restrict__ evals,
                                                 const double* __restrict__ evecs, double* __restrict__ t,
                                                 double rel_eps, const int* __restrict__ skip_flags) {
  const uint32_t b = blockIdx.x;
  if (skip_flags[b]) return;
  const uint32_t l = probs[b].l;
  const double* ev = evals + (size_t)b * 32;
  const double lmax = ev[0];
  for (int i = threadIdx.x; i < (int)(l * l); i += blockDim.x) {
    const double lam = ev[i % l];
    t[(size_t)b * 1024 + i] = (lam > rel_eps * lmax && lam > 0.0) ? evecs[(size_t)b * 1024 + i] / sqrt(lam) : 0.0;
  }
}

int launch_make_orth_transform_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, double rel_eps,
                                     const DenseBatchWs& ws) {
  make_orth_transform_batch_kernel<<<n_probs, 128, 0, c->stream>>>(d_probs, ws.evals, ws.evecs, ws.T, rel_eps, ws.flags);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

__global__ void rotation_batch_kernel(const DenseProb* __restrict__ probs, const uint32_t* __restrict__ l2s,
                                      const double* __restrict__ evecs, double* __restrict__ t) {
  const uint32_t b = blockIdx.x;
  const uint32_t l = probs[b].l, k = l2s[b];
  for (int i = threadIdx.x; i < (int)(l * k); i += blockDim.x) {
    const int r = i / k, j = i % k;
    t[(size_t)b * 1024 + i] = evecs[(size_t)b * 1024 + r * l + j];
  }
}

int launch_rotation_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, const uint32_t* d_l2,
                          const DenseBatchWs& ws) {
  rotation_batch_kernel<<<n_probs, 128, 0, c->stream>>>(d_probs, d_l2, ws.evecs, ws.T);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// grid = (row tiles, problems); same tiling as apply_right_kernel (128-row f64 tiles, row pairs, column pairs)
template <int NOWN>
__global__ void __launch_bounds__(256) apply_right_batch_kernel(const float* __restrict__ base, uint32_t ld,
                                                                const DenseProb* __restrict__ probs,
                                                                const double* __restrict__ t_all,
                                                                const uint32_t* __restrict__ l2s,
                                                                float* __restrict__ out_base,
                                                                const uint64_t* __restrict__ out_offs, uint32_t ldo) {
  static_assert(NOWN % 2 == 0, "columns are owned in pairs");
  extern __shared__ __align__(16) double sm[];
  const DenseProb pb = probs[blockIdx.y];
  const uint32_t l = pb.l;
  const uint32_t l2 = l2s ? l2s[blockIdx.y] : l;
  const uint64_t n = pb.rows;
  const float* y = base + pb.off;
  float* out = out_base + (out_offs ? out_offs[blockIdx.y] : pb.off);
  const double* t = t_all + (size_t)blockIdx.y * 1024;
  const int l2p = NOWN * 4;
  double* ts = sm;                                             // [32][l2p]
  double* tile = sm + 32 * l2p;                                // [AR_ROWS][33]
  for (int i = threadIdx.x; i < (int)(l * l2p); i += 256) {
    const int cc = i / l2p, c2 = i % l2p;
    ts[i] = ((uint32_t)c2 < l2) ? t[cc * l2 + c2] : 0.0;
  }
  const int lp = 33;
  const uint64_t ntiles = (n + AR_ROWS - 1) / AR_ROWS;
  const bool vec2 = (ldo & 1u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
  // warp = row, lane = column; for l <= 32 the next tile's loads are in flight while the current tile is multiplied
  float stg[AR_ROWS / 8];
  auto fetch = [&](uint64_t r0, uint32_t c0) {
    const uint32_t cc = c0 + (threadIdx.x & 31);
#pragma unroll
    for (int q = 0; q < AR_ROWS / 8; ++q) {
      const uint64_t r = r0 + (threadIdx.x >> 5) + 8 * q;
      stg[q] = (r < n && cc < l) ? __ldg(y + r * ld + cc) : 0.0f;
    }
  };
  if (l <= 32 && blockIdx.x < ntiles) fetch((uint64_t)blockIdx.x * AR_ROWS, 0);
  for (uint64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const uint64_t r0 = tix * AR_ROWS;
    __syncthreads();
    if (l <= 32) {      // registers prefetched during the previous tile's multiply -> f64 tile
      const uint32_t cc = threadIdx.x & 31;
      if (cc < l) {
#pragma unroll
        for (int q = 0; q < AR_ROWS / 8; ++q) tile[((threadIdx.x >> 5) + 8 * q) * lp + cc] = (double)stg[q];
      }
    } else {
      for (uint32_t c0 = 0; c0 < l; c0 += 32) {
        fetch(r0, c0);
        const uint32_t cc = c0 + (threadIdx.x & 31);
        if (cc < l) {
#pragma unroll
          for (int q = 0; q < AR_ROWS / 8; ++q) tile[((threadIdx.x >> 5) + 8 * q) * lp + cc] = (double)stg[q];
        }
      }
    }
    __syncthreads();
    if (l <= 32 && tix + gridDim.x < ntiles) fetch((tix + gridDim.x) * AR_ROWS, 0);
    const int rr = threadIdx.x >> 2, cg = threadIdx.x & 3;
    double acc0[NOWN], acc1[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; ++j) acc0[j] = acc1[j] = 0.0;
    const double* y0 = tile + rr * lp;
    const double* y1 = tile + (rr + 64) * lp;
    for (uint32_t cc = 0; cc < l; ++cc) {
      const double a0 = y0[cc], a1 = y1[cc];
      const double* trow = ts + cc * l2p + 2 * cg;
#pragma unroll
      for (int j = 0; j < NOWN / 2; ++j) {
        const double2 tv = *reinterpret_cast<const double2*>(trow + 8 * j);
        acc0[2 * j] = fma(a0, tv.x, acc0[2 * j]);
        acc0[2 * j + 1] = fma(a0, tv.y, acc0[2 * j + 1]);
        acc1[2 * j] = fma(a1, tv.x, acc1[2 * j]);
        acc1[2 * j + 1] = fma(a1, tv.y, acc1[2 * j + 1]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t r = r0 + rr + 64 * h;
      if (r >= n) continue;
      const double* acc = h ? acc1 : acc0;
      float* orow = out + r * ldo;
#pragma unroll
      for (int j = 0; j < NOWN / 2; ++j) {
        const uint32_t c0 = 2 * cg + 8 * j;
        if (vec2 && c0 + 1 < l2) {
          *reinterpret_cast<float2*>(orow + c0) = make_float2((float)acc[2 * j], (float)acc[2 * j + 1]);
        } else {
          if (c0 < l2) orow[c0] = (float)acc[2 * j];
          if (c0 + 1 < l2) orow[c0 + 1] = (float)acc[2 * j + 1];
        }
      }
    }
  }
}

template <int NOWN>
static int run_apply_right_batch(gpca_ctx* c, const float* d_base, uint32_t ld, const DenseProb* d_probs,
                                 uint32_t n_probs, uint64_t max_rows, const double* d_t, const uint32_t* d_l2,
                                 float* d_out, const uint64_t* d_out_offs, uint32_t ldo) {
  const size_t smem = ((size_t)32 * NOWN * 4 + (size_t)AR_ROWS * 33) * sizeof(double);
  if (smem > 48 * 1024)
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(apply_right_batch_kernel<NOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
  uint64_t gx = (max_rows + AR_ROWS - 1) / AR_ROWS;
  const uint64_t cap = ((uint64_t)c->sm_count * 16 + n_probs - 1) / n_probs;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, n_probs);
  apply_right_batch_kernel<NOWN><<<grid, 256, smem, c->stream>>>(d_base, ld, d_probs, d_t, d_l2, d_out, d_out_offs, ldo);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_apply_right_batch(gpca_ctx* c, const float* d_base, uint32_t ld, const DenseProb* d_probs,
                             uint32_t n_probs, uint64_t max_rows, const double* d_t, const uint32_t* d_l2,
                             uint32_t max_l2, float* d_out, const uint64_t* d_out_offs, uint32_t ldo) {
  if (!n_probs || !max_rows || !max_l2) return GPCA_OK;
  if (max_l2 > 32) {
    c->set_error("apply_right_batch: l2 must be <= 32");
    return GPCA_ERR_INVALID;
  }
  const uint32_t need = (max_l2 + 3) / 4;
  if (need <= 2) return run_apply_right_batch<2>(c, d_base, ld, d_probs, n_probs, max_rows, d_t, d_l2, d_out, d_out_offs, ldo);
  if (need <= 6) return run_apply_right_batch<6>(c, d_base, ld, d_probs, n_probs, max_rows, d_t, d_l2, d_out, d_out_offs, ldo);
  return run_apply_right_batch<8>(c, d_base, ld, d_probs, n_probs, max_rows, d_t, d_l2, d_out, d_out_offs, ldo);
}

int orthonormalize_batch(gpca_ctx* c, float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                         uint64_t max_rows, uint32_t max_l, const DenseBatchWs& ws) {
  for (int rep = 0; rep < 2; ++rep) {
    int np = 1;
    GPCA_TRY(launch_gram_batch(c, d_base, ld, d_probs, n_probs, max_rows, &np));
    const double eps = rep == 0 ? 1e-11 : 1e-13;
    GPCA_TRY(launch_chol_orth_batch(c, np, d_probs, n_probs, eps, ws));
    GPCA_TRY(launch_jacobi_eigh_batch(c, d_probs, n_probs, ws, true));
    GPCA_TRY(launch_make_orth_transform_batch(c, d_probs, n_probs, eps, ws));
    GPCA_TRY(launch_apply_right_batch(c, d_base, ld, d_probs, n_probs, max_rows, ws.T, nullptr, max_l, d_base, nullptr, ld));
  }
  return GPCA_OK;
}
