// kernels_sketch.cu -- the sketch pass   out = a (.) (C * (f (.) Bin)) - b (x) (e^T Bin)
// over the resident 2-bit matrix C (dosage codes), i.e. one half-step of the randomized range
// finder on the standardized matrix without ever materialising it:
//   snp side    (rows = SNPs,    K = samples):  a = 1/sd, b = mean/sd, f = e = 1
//   sample side (rows = samples, K = SNPs)   :  a = b = 1,  f = 1/sd,   e = mean/sd
// Replaces get_standardized_snp_sample_block (src/prepare.rs:1884-2016, f32 standardise) plus the
// GEMMs efficient_pca runs on those blocks.  Missing calls (code 3) contribute 0 after
// standardisation (mean imputation by mask; the reference errors instead, prepare.rs:1906-1912).
//
// This file holds: operand preparation, the SIMT fp32 engine (engine 0; also the fallback for
// shapes the tcgen05 engine does not take), split-K reduction + epilogue, missing correction.
#include "kernels.cuh"
#include "sketch_tc.cuh"

#define KLAUNCH_CHECK(c)                    \
  do {                                      \
    (c)->launches++;                        \
    GPCA_CUDA_TRY((c), cudaGetLastError()); \
  } while (0)

// ------------------------------------------------------------------------------------------
// prep: Bp[k][c] = f_k * Bin[k][c] (zero padded to [Kpad x NC]);  cpart[block][c] = sum_k e_k Bin[k][c] (f64)
template <int NC>
__global__ void __launch_bounds__(256) prep_b_kernel(const float* __restrict__ bin, uint64_t K, uint64_t Kpad,
                                                     uint32_t l, uint32_t ld, const float* __restrict__ f,
                                                     const float* __restrict__ e, float* __restrict__ bp,
                                                     double* __restrict__ cpart) {
  constexpr int RPB = 256 / NC;  // rows per block iteration
  __shared__ double red[256];
  const int cidx = threadIdx.x % NC;
  const int rr = threadIdx.x / NC;
  double acc = 0.0;
  for (uint64_t k = (uint64_t)blockIdx.x * RPB + rr; k < Kpad; k += (uint64_t)gridDim.x * RPB) {
    float v = 0.0f, w = 0.0f;
    if (k < K && (uint32_t)cidx < l) {
      const float x = bin[k * ld + cidx];
      v = f ? x * f[k] : x;
      w = e ? x * e[k] : x;
    }
    bp[k * NC + cidx] = v;
    acc += (double)w;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rr == 0) {
    double s = 0.0;
    for (int i = 0; i < RPB; ++i) s += red[i * NC + cidx];
    cpart[(uint64_t)blockIdx.x * NC + cidx] = s;
  }
}

template <int NC>
__global__ void cvec_reduce_kernel(const double* __restrict__ cpart, int nparts, float* __restrict__ cvec) {
  const int cidx = threadIdx.x;
  if (cidx < NC) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += cpart[(uint64_t)p * NC + cidx];
    cvec[cidx] = (float)s;
  }
}

// ------------------------------------------------------------------------------------------
// SIMT engine: CTA tile 128 rows x NC columns, K chunk 32, 4 x (NC/8) register tile per thread.
template <int NC>
__global__ void __launch_bounds__(256) sketch_simt_kernel(const uint8_t* __restrict__ g, size_t pitch, uint64_t rows,
                                                          uint64_t kchunks_total, uint32_t kchunks_per_split,
                                                          const float* __restrict__ bp, const float* __restrict__ a,
                                                          const float* __restrict__ b, const float* __restrict__ cvec,
                                                          float* __restrict__ out, uint32_t ldo, uint32_t l,
                                                          float* __restrict__ partial) {
  constexpr int CPT = NC / 8;  // columns per thread
  __shared__ __align__(16) float As[32][132];
  __shared__ __align__(16) float Bs[32][NC];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const uint64_t row0 = (uint64_t)blockIdx.x * 128;
  const uint64_t kc_begin = (uint64_t)blockIdx.y * kchunks_per_split;
  uint64_t kc_end = kc_begin + kchunks_per_split;
  if (kc_end > kchunks_total) kc_end = kchunks_total;
  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.0f;

  const int drow = threadIdx.x & 127, dhalf = threadIdx.x >> 7;
  const bool drow_ok = (row0 + drow) < rows;
  const uint8_t* grow = g + (row0 + (drow_ok ? drow : 0)) * pitch;
  for (uint64_t kc = kc_begin; kc < kc_end; ++kc) {
    // decode 16 fields per thread
    uint32_t w = 0;
    if (drow_ok) w = *reinterpret_cast<const uint32_t*>(grow + kc * 8 + dhalf * 4);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t code = (w >> (2 * j)) & 3u;
      As[dhalf * 16 + j][drow] = (code == 3u) ? 0.0f : (float)code;
    }
    // stage B chunk
    {
      const float4* src = reinterpret_cast<const float4*>(bp + kc * 32 * NC);
      float4* dst = reinterpret_cast<float4*>(&Bs[0][0]);
#pragma unroll
      for (int i = 0; i < NC / 32; ++i) dst[threadIdx.x + i * 256] = src[threadIdx.x + i * 256];
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float avv[4] = {av.x, av.y, av.z, av.w};
      float bv[CPT];
#pragma unroll
      for (int q = 0; q < CPT / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tx * CPT + q * 4]);
        bv[q * 4 + 0] = t.x;
        bv[q * 4 + 1] = t.y;
        bv[q * 4 + 2] = t.z;
        bv[q * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(avv[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t r = row0 + ty * 4 + i;
    if (r >= rows) continue;
    if (partial) {
      float* p = partial + ((uint64_t)blockIdx.y * rows + r) * NC + tx * CPT;
#pragma unroll
      for (int j = 0; j < CPT; ++j) p[j] = acc[i][j];
    } else {
      const float ar = a ? a[r] : 1.0f, br = b ? b[r] : 1.0f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t cc = tx * CPT + j;
        if (cc < l) out[r * ldo + cc] = ar * acc[i][j] - br * cvec[cc];
      }
    }
  }
}

// split-K reduction + epilogue (deterministic order)
template <int NC>
__global__ void sketch_reduce_kernel(const float* __restrict__ partial, int nsplit, uint64_t rows,
                                     const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ cvec, float* __restrict__ out, uint32_t ldo,
                                     uint32_t l, float acc_scale) {
  const uint64_t total = rows * NC;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / NC;
    const uint32_t cc = (uint32_t)(t % NC);
    if (cc >= l) continue;
    float s = 0.0f;
    for (int p = 0; p < nsplit; ++p) s += partial[((uint64_t)p * rows + r) * NC + cc];
    const float ar = a ? a[r] : 1.0f, br = b ? b[r] : 1.0f;
    out[r * ldo + cc] = ar * (s * acc_scale) - br * cvec[cc];
  }
}

// ------------------------------------------------------------------------------------------
// Missing-call correction (warp per row):
//   out[r,:] += sum_{k : code(r,k)==3} (b_r e_k - a_r miss_val f_k) Bin[k,:]
// miss_val = what the engine's product used for a missing field (0: SIMT engine, 3: tcgen05 engine).
__global__ void __launch_bounds__(256) missing_fix_kernel(const uint8_t* __restrict__ g, size_t pitch, uint64_t rows,
                                                          uint64_t K, const float* __restrict__ bin, uint32_t l,
                                                          uint32_t ld, const float* __restrict__ e,
                                                          const float* __restrict__ f, const float* __restrict__ a,
                                                          const float* __restrict__ b, float miss_val,
                                                          float* __restrict__ out, uint32_t ldo) {
  const uint64_t warp_id = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t words = (K + 15) / 16;
  for (uint64_t r = warp_id; r < rows; r += nwarps) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(g + r * pitch);
    const float br = b ? b[r] : 1.0f;
    const float am = (a ? a[r] : 1.0f) * miss_val;
    float c0 = 0.0f, c1 = 0.0f;
    bool any = false;
    for (uint64_t w0 = 0; w0 < words; w0 += 32) {
      const uint64_t wi = w0 + lane;
      uint32_t w = (wi < words) ? p[wi] : 0u;
      if (wi + 1 == words && (K & 15)) w &= (1u << (2 * (int)(K & 15))) - 1u;   // a view's last word may hold a neighbour's fields
      uint32_t miss = w & (w >> 1) & 0x55555555u;
      unsigned ballot = __ballot_sync(0xffffffffu, miss != 0u);
      while (ballot) {
        const int src = __ffs(ballot) - 1;
        ballot &= ballot - 1;
        uint32_t m = __shfl_sync(0xffffffffu, miss, src);
        any = true;
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t k = (w0 + src) * 16 + (bit >> 1);
          const float wk = br * (e ? e[k] : 1.0f) - am * (f ? f[k] : 1.0f);
          if ((uint32_t)lane < l) c0 = fmaf(wk, bin[k * ld + lane], c0);
          if ((uint32_t)(lane + 32) < l) c1 = fmaf(wk, bin[k * ld + lane + 32], c1);
        }
      }
    }
    if (any) {
      if ((uint32_t)lane < l) out[r * ldo + lane] += c0;
      if ((uint32_t)(lane + 32) < l) out[r * ldo + lane + 32] += c1;
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int NC>
static int run_simt(gpca_ctx* c, const SketchProblem& p, uint64_t Kpad) {
  const uint64_t rows = p.G.rows;
  const uint64_t kchunks = Kpad / 32;
  const uint64_t row_tiles = (rows + 127) / 128;
  // split K so that at least ~2 CTAs per SM exist, each split >= 16 chunks
  uint64_t nsplit = 1;
  if (row_tiles < (uint64_t)c->sm_count * 2) {
    nsplit = ((uint64_t)c->sm_count * 2 + row_tiles - 1) / row_tiles;
    const uint64_t max_split = (kchunks + 15) / 16;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
  }
  uint32_t per = (uint32_t)((kchunks + nsplit - 1) / nsplit);
  nsplit = (kchunks + per - 1) / per;
  float* partial = nullptr;
  if (nsplit > 1) {
    GPCA_CUDA_TRY(c, c->ws_partial.alloc(nsplit * rows * NC));
    partial = c->ws_partial.p;
  }
  dim3 grid((unsigned)row_tiles, (unsigned)nsplit);
  KernelTimer kt(c);
  sketch_simt_kernel<NC><<<grid, 256, 0, c->stream>>>(p.G.p, p.G.pitch, rows, kchunks, per, c->ws_bprep.p, p.a, p.b,
                                                      c->ws_cvec.p, p.out, p.ldo, p.l, partial);
  kt.end();
  KLAUNCH_CHECK(c);
  if (nsplit > 1) {
    const uint64_t total = rows * NC;
    const uint64_t blocks = (total + 255) / 256;
    const int g2 = (int)(blocks < (uint64_t)c->sm_count * 8 ? blocks : (uint64_t)c->sm_count * 8);
    sketch_reduce_kernel<NC><<<g2, 256, 0, c->stream>>>(partial, (int)nsplit, rows, p.a, p.b, c->ws_cvec.p, p.out,
                                                        p.ldo, p.l, 1.0f);
    KLAUNCH_CHECK(c);
  }
  return GPCA_OK;
}

template <int NC>
static int run_prep(gpca_ctx* c, const SketchProblem& p, uint64_t Kpad) {
  const uint64_t K = p.G.cols;
  GPCA_CUDA_TRY(c, c->ws_bprep.alloc(Kpad * NC));
  int nblocks = (int)((Kpad * NC + 255) / 256);
  if (nblocks > c->sm_count * 8) nblocks = c->sm_count * 8;
  if (nblocks < 1) nblocks = 1;
  GPCA_CUDA_TRY(c, c->ws_cpart.alloc((size_t)nblocks * NC));
  GPCA_CUDA_TRY(c, c->ws_cvec.alloc(64));
  prep_b_kernel<NC><<<nblocks, 256, 0, c->stream>>>(p.Bin, K, Kpad, p.l, p.ld, p.f, p.e, c->ws_bprep.p, c->ws_cpart.p);
  KLAUNCH_CHECK(c);
  cvec_reduce_kernel<NC><<<1, 64, 0, c->stream>>>(c->ws_cpart.p, nblocks, c->ws_cvec.p);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

int launch_sketch(gpca_ctx* c, const SketchProblem& p) {
  if (p.l == 0 || p.l > 64) {
    c->set_error("sketch: l must be in 1..64");
    return GPCA_ERR_INVALID;
  }
  if (p.G.rows == 0 || p.G.cols == 0) return GPCA_OK;
  if (p.gen) {
    // generated operand: only the integer engine quantises it on the fly (and only without missing calls: the
    // correction kernel reads Bin); everything else gets the matrix written into the caller's scratch first
    const bool fused = c->engine == 2 && sketch_i8_supported(c, p) && !c->any_missing && p.gen_amax > 0.0f &&
                       !getenv("GPCA_DEBUG_NO_GEN_FUSE");
    if (!fused) {
      GPCA_TRY(launch_gaussian(c, const_cast<float*>(p.Bin), p.G.cols, p.l, p.ld, p.gen_seed, p.gen_stream, p.gen_row0));
      SketchProblem q = p;
      q.gen = false;
      q.use_stats = false;
      return launch_sketch(c, q);
    }
  }
  int rc = GPCA_ERR_INVALID;
  bool done = false;
  if (c->engine == 2 && sketch_i8_supported(c, p)) {
    rc = launch_sketch_i8(c, p);
    done = true;
    c->last_engine = 2;
  } else if (c->engine >= 1 && sketch_tc_supported(c, p)) {
    rc = launch_sketch_tc(c, p);
    done = true;
    c->last_engine = 1;
  }
  if (!done) {
    c->last_engine = 0;
    const uint64_t Kpad = round_up(p.G.cols, 32);
    if (p.l <= 32) {
      GPCA_TRY(run_prep<32>(c, p, Kpad));
      rc = run_simt<32>(c, p, Kpad);
    } else {
      GPCA_TRY(run_prep<64>(c, p, Kpad));
      rc = run_simt<64>(c, p, Kpad);
    }
  }
  if (rc != GPCA_OK) return rc;
  if (c->any_missing) {
    const uint64_t need = (p.G.rows * 32 + 255) / 256;
    const int grid = (int)(need < (uint64_t)c->sm_count * 8 ? need : (uint64_t)c->sm_count * 8);
    missing_fix_kernel<<<grid, 256, 0, c->stream>>>(p.G.p, p.G.pitch, p.G.rows, p.G.cols, p.Bin, p.l, p.ld, p.e, p.f,
                                                    p.a, p.b, done ? 3.0f : 0.0f, p.out, p.ldo);
    KLAUNCH_CHECK(c);
  }
  return GPCA_OK;
}
