// eigensnp.cu -- EigenSNP driver: replaces EigenSNPCoreAlgorithm::compute_pca(&accessor, &blocks)
// (src/main.rs:359-366; config src/main.rs:311-327; effective defaults src/main.rs:545-588).
// The algorithm lives in the external efficient_pca crate (parity unpinned); the stage order here is the
// one restated in oracle/pca.py::eigensnp, which the tests check this driver against:
//   1. sample subset N_s = clamp(subset_factor*N, min, max)
//   2. per LD block: randomized SVD of the standardized block on the subset -> local basis U_p [M_p x c_p]
//   3. condensed features C_p = U_p^T X_p for all N samples, stacked, row-standardised
//   4. global randomized SVD of the condensed matrix -> initial sample-side vectors V [N x k]
//   5. refine passes: L = orth(S V); Sc = S^T L; small eigensolve -> V, loadings, singular values
// Every product with genotypes is a sketch pass on the resident 2-bit matrices (no accessor round trips,
// no f32 strips: src/prepare.rs:1839-2022 is what this makes unnecessary); the products with the dense condensed
// matrix run on the split-bf16 tcgen05 engine of dense_tc.cu, with the column standardisation folded in.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "dense_tc.cuh"
#include "driver_util.cuh"
#include "parallel_for.h"
#include "philox.cuh"

static int fail(gpca_ctx* c, int code, const std::string& msg) {
  c->set_error(msg);
  return code;
}

#define KCHECK(c)                           \
  do {                                      \
    (c)->launches++;                        \
    GPCA_CUDA_TRY((c), cudaGetLastError()); \
  } while (0)

namespace {
constexpr uint32_t STREAM_GLOBAL = 2;
constexpr uint32_t STREAM_SUBSET = 100;
constexpr uint32_t STREAM_LOCAL0 = 1000;

// Sample-major slot-ordered copy when every LD block is a run of consecutive PcaSnpIds (the normal case: blocks are
// genomic intervals, src/prepare.rs:1424-1563): Et[n, 64q .. 64q+63] = Gt[n, first[q] .. first[q] + count[q] - 1], a
// per-row shifted copy of 2-bit fields -- one 16-byte chunk per thread, two aligned 16-byte loads and a funnel shift.
// (The general path gathers SNP-major rows and transposes the whole matrix: 16 ms for a 10.9 GB shard against ~4 ms.)
__global__ void shift_fields_kernel(const uint8_t* __restrict__ src, size_t src_pitch, const int64_t* __restrict__ first,
                                    const uint32_t* __restrict__ count, uint8_t* __restrict__ dst, size_t dst_pitch,
                                    uint64_t n_rows, uint64_t n_chunks) {
  const uint64_t cpr = dst_pitch / 16;
  const uint64_t total = n_rows * cpr;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / cpr, ci = t - r * cpr;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (ci < n_chunks) {
      const int64_t f0 = first[ci];
      const uint32_t cnt = count[ci];
      if (f0 >= 0 && cnt) {
        const uint64_t byte0 = ((uint64_t)f0 >> 6) << 4;          // aligned 16-byte word holding field f0
        const uint32_t sh = ((uint32_t)f0 & 63u) * 2u;            // bit offset inside it
        const uint8_t* row = src + r * src_pitch;
        const uint4 a = ldg_nc_v4(row + byte0);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (sh && byte0 + 32 <= src_pitch) b = ldg_nc_v4(row + byte0 + 16);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const uint32_t ws = sh >> 5, bs = sh & 31u;
        uint32_t v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {      // v[i] = w[i + ws] without dynamic indexing
          v[i] = (ws == 0) ? w[i] : (ws == 1) ? w[i + 1] : (ws == 2) ? w[i + 2] : w[i + 3];
        }
        o.x = __funnelshift_r(v[0], v[1], bs);
        o.y = __funnelshift_r(v[1], v[2], bs);
        o.z = __funnelshift_r(v[2], v[3], bs);
        o.w = __funnelshift_r(v[3], v[4], bs);
        if (cnt < 64) {                    // fields beyond the block are padding: dosage code 0
          const uint32_t keep_bits = cnt * 2u;
          uint32_t* ow = &o.x;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t lo = 32u * i;
            if (keep_bits <= lo) ow[i] = 0u;
            else if (keep_bits < lo + 32u) ow[i] &= (1u << (keep_bits - lo)) - 1u;
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(dst + r * dst_pitch + ci * 16) = o;
  }
}

// Block-diagonal operand of the grouped condensed-feature pass: dst [positions x 32] (zeroed by the caller); a
// position of block p gets U_p's row (src [positions x ld_src]) in columns col0(p) .. col0(p) + c_p - 1; positions
// outside every block (padding slots) stay zero.
__global__ void block_diag_operand_kernel(const float* __restrict__ src, uint32_t ld_src,
                                          const uint32_t* __restrict__ blk_of_pos, const uint32_t* __restrict__ col0,
                                          const uint32_t* __restrict__ cpn, uint64_t n_pos, float* __restrict__ dst) {
  const uint64_t total = n_pos * ld_src;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = t / ld_src;
    const uint32_t j = (uint32_t)(t - s * ld_src);
    const uint32_t b = blk_of_pos[s];
    if (b != 0xffffffffu && j < cpn[b]) dst[s * 32 + col0[b] + j] = src[t];
  }
}

// Gaussian rows keyed by an explicit 64-bit key per row (condensed-feature test matrix)
__global__ void gaussian_keyed_kernel(float* __restrict__ out, const uint64_t* __restrict__ keys, uint64_t rows,
                                      uint32_t cols, uint64_t seed, uint32_t stream) {
  const uint64_t total = rows * cols;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / cols;
    out[t] = philox_normal(seed, stream, keys[r], (uint32_t)(t - r * cols));
  }
}

// per-column mean and 1/sd (ddof = 1) of X [n x r] row-major with row stride ld, f64 accumulation.  The standardised
// matrix z = (x - mean) / sd is never written: the products with it take mean / sd as operand scales (dense_tc.cu).
__global__ void __launch_bounds__(256) col_moments_kernel(const float* __restrict__ x, uint64_t n, uint32_t r, uint32_t ld,
                                                          uint64_t rows_per_cta, double* __restrict__ part) {
  // grid.x = row chunks, grid.y = column tiles of 256
  const uint32_t col = blockIdx.y * 256 + threadIdx.x;
  const uint64_t r0 = blockIdx.x * rows_per_cta;
  const uint64_t r1 = (r0 + rows_per_cta < n) ? r0 + rows_per_cta : n;
  double s = 0.0, ss = 0.0;
  if (col < r)
    for (uint64_t i = r0; i < r1; ++i) {
      const double v = (double)x[i * ld + col];
      s += v;
      ss += v * v;
    }
  if (col < r) {
    part[((uint64_t)blockIdx.x * r + col) * 2 + 0] = s;
    part[((uint64_t)blockIdx.x * r + col) * 2 + 1] = ss;
  }
}
__global__ void col_finalize_kernel(const double* __restrict__ part, int nparts, uint64_t n, uint32_t r,
                                    float* __restrict__ mean, float* __restrict__ inv_sd, float* __restrict__ mean_inv_sd) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= r) return;
  double s = 0.0, ss = 0.0;
  for (int q = 0; q < nparts; ++q) {
    s += part[((uint64_t)q * r + col) * 2 + 0];
    ss += part[((uint64_t)q * r + col) * 2 + 1];
  }
  const double m = s / (double)n;
  double var = (n > 1) ? (ss - (double)n * m * m) / (double)(n - 1) : 0.0;
  if (var < 0.0) var = 0.0;
  const double sd = sqrt(var);
  const double inv = (sd > 1e-12) ? 1.0 / sd : 0.0;
  mean[col] = (float)m;
  inv_sd[col] = (float)inv;
  mean_inv_sd[col] = (float)(m * inv);
}

// out[n x k] = in[n x k] * diag(scale[k]) in place (f64 scale vector on device, sqrt applied)
__global__ void scale_cols_sqrt_kernel(float* __restrict__ x, uint64_t n, uint32_t k, const double* __restrict__ lam) {
  const uint64_t total = n * k;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const double l = lam[t % k];
    x[t] = (float)((double)x[t] * (l > 0.0 ? sqrt(l) : 0.0));
  }
}

// out[id_of_slot[s]][j] = +-in[s][j]  (pad slots skipped)
__global__ void scatter_loadings_kernel(const float* __restrict__ in, const int64_t* __restrict__ id_of_slot,
                                        uint64_t n_slots, uint32_t k, const int* __restrict__ flags,
                                        float* __restrict__ out) {
  const uint64_t total = n_slots * k;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t sidx = t / k;
    const uint32_t j = (uint32_t)(t - sidx * k);
    const int64_t id = id_of_slot[sidx];
    if (id < 0) continue;
    const float v = in[t];
    out[(uint64_t)id * k + j] = flags[j] ? -v : v;
  }
}

}  // namespace

extern "C" uint64_t gpca_eigensnp_workspace_bytes(uint64_t N, uint64_t D, uint64_t n_blocks, const gpca_eigensnp_cfg* cfg) {
  gpca_eigensnp_cfg d;
  if (!cfg) {
    gpca_eigensnp_default_cfg(&d);
    cfg = &d;
  }
  uint64_t Ns = (uint64_t)(cfg->subset_factor * (double)N);
  Ns = std::min<uint64_t>(N, std::max<uint64_t>(cfg->min_subset_size, std::min<uint64_t>(Ns, cfg->max_subset_size)));
  const uint64_t cpb = cfg->components_per_ld_block, LD = cpb + cfg->local_oversampling;
  const uint64_t R = n_blocks * cpb, lg = cfg->target_num_global_pcs + cfg->global_oversampling;
  uint64_t b = N * R * 4;                                                  // condensed features
  b += Ns * round_up((D + 3) / 4 + 64, 128) + D * round_up((Ns + 3) / 4, 128);   // subset copies, both orientations
  b += n_blocks * Ns * LD * 4 + D * (LD + cpb + 32) * 4;                   // per-block iterates, bases, grouped operand
  b += N * lg * 4 * 4 + R * lg * 4 * 2 + D * lg * 4;                       // global iterates, scores, loadings
  b += n_blocks * round_up(Ns, 256) * 64 + (D + 64 * n_blocks + N) * 64 * 2;   // operand images of the per-block passes
  b += n_blocks * (3 * 1024 + 64) * 8 * 2 + (D + N) * 32 * 4 * 4;          // per-block small matrices, split-K partials
  b += 6ull << 30;                                                         // Gram partials, staging, allocator slack
  return b;
}

extern "C" int gpca_eigensnp(gpca_ctx* c, const gpca_eigensnp_cfg* cfg, const uint64_t* block_offsets,
                             uint64_t n_blocks, const uint64_t* block_snp_ids, float* scores, double* eigenvalues,
                             float* loadings, uint32_t* k_out) {
  if (!c) return GPCA_ERR_INVALID;
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->D == 0) return fail(c, GPCA_ERR_INVALID, "gpca_set_pca_snps has not been called");
  if (!cfg || !block_offsets || !block_snp_ids || n_blocks == 0)
    return fail(c, GPCA_ERR_INVALID, "No SNPs mapped to LD blocks or all resulting blocks were empty.");  // prepare.rs:1031
  const uint64_t N = c->N, D = c->D;
  if (N < 2) return fail(c, GPCA_ERR_INVALID, "EigenSNP requires at least 2 samples");
  // (the reference's accessor errors on any missing call, prepare.rs:1906-1912; here missing calls are mean-imputed)
  const uint32_t k_req = cfg->target_num_global_pcs;
  const uint32_t cpb_max = cfg->components_per_ld_block;
  if (k_req == 0 || cpb_max == 0) return fail(c, GPCA_ERR_INVALID, "k_global and components_per_block must be > 0");
  if (cpb_max + cfg->local_oversampling > 64 || k_req + cfg->global_oversampling > 64)
    return fail(c, GPCA_ERR_INVALID, "components + oversampling must be <= 64 in this build");
  const uint64_t seed = cfg->random_seed;
  // optional stage timing (GPCA_TRACE=1): the reference prints a stage table too (src/main.rs:437-442).  With
  // cfg->collect_diagnostics (--eigensnp-collect-diagnostics, src/main.rs:411-430) the stage times and the shape of the
  // run are kept as a JSON document (gpca_eigensnp_diagnostics); the stream is then synchronised at every stage.
  const bool trace = getenv("GPCA_TRACE") != nullptr;
  const bool diag = cfg->collect_diagnostics != 0;
  std::vector<std::pair<std::string, double>> diag_stages;
  const auto t_call = std::chrono::steady_clock::now();
  auto t_last = t_call;
  auto stage = [&](const char* name) {
    if (!trace && !diag) return;
    if (diag || strcmp(getenv("GPCA_TRACE"), "2") != 0) cudaStreamSynchronize(c->stream);   // "2": host-side times only
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - t_last).count();
    if (trace) fprintf(stderr, "[gpca_eigensnp] %-28s %9.2f ms\n", name, ms);
    if (diag) {
      std::string nm(name);
      nm.erase(0, nm.find_first_not_of(' '));
      diag_stages.push_back({nm, ms});
    }
    t_last = now;
  };
  const uint64_t launches0 = c->launches, collectives0 = c->collectives;
  const double sk_bytes0 = c->sk_bytes;
  const uint64_t sk_passes0 = c->sk_passes;

  c->es_pool.reset();
  // ---- subset of samples for the local bases ----------------------------------------------------------
  uint64_t Ns = (uint64_t)(cfg->subset_factor * (double)N);
  Ns = std::max<uint64_t>(cfg->min_subset_size, std::min<uint64_t>(Ns, cfg->max_subset_size));
  Ns = std::min<uint64_t>(Ns, N);
  // (kept in the context: the subset depends only on N, N_s and the seed; the Philox keys are drawn on all host
  //  threads and the selection is an nth_element -- ~8 ms single-threaded at N = 500,000)
  std::vector<int64_t>& sub = c->es_subset;
  if (!(c->es_subset_n == N && c->es_subset_seed == seed && sub.size() == Ns)) {
    sub.assign(Ns, 0);
    if (Ns == N) {
      for (uint64_t i = 0; i < N; ++i) sub[i] = (int64_t)i;
    } else {
      std::vector<std::pair<uint32_t, uint64_t>> keys(N);
      parallel_for(N, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
          const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, STREAM_SUBSET, (uint32_t)seed,
                                          (uint32_t)(seed >> 32));
          keys[i] = {r.x, i};
        }
      }, 1u << 14);
      // the N_s smallest (key, index) pairs -- ties by index, the oracle's stable order; a selection is enough because
      // the subset is then put in index order
      std::nth_element(keys.begin(), keys.begin() + Ns, keys.end());
      for (uint64_t i = 0; i < Ns; ++i) sub[i] = (int64_t)keys[i].second;
      std::sort(sub.begin(), sub.end());
    }
    c->es_subset_n = N;
    c->es_subset_seed = seed;
  }

  // ---- positions: where a block's SNPs sit in the matrices the per-block passes read ---------------------------------
  // Two layouts.  ID ORDER: when every LD block is a run of consecutive PcaSnpIds (blocks are genomic intervals,
  // src/prepare.rs:1424-1563 -- the normal case) and the blocks cover every PCA SNP, a block's positions are its
  // PcaSnpIds and every pass runs on the resident matrices themselves (Gs / Gt and their 1/sd, mean/sd vectors): no
  // slot-ordered copy of the packed matrix exists (at 500,000 x 700,000 each copy is 87.5 GB).  A block then starts at
  // an arbitrary field of a sample-major row; a K range starts at the 64-field boundary below it and the operand image
  // gets zero rows for the fields in front of the block (SketchBatchBlock::kskip).  SLOT ORDER (any block list, and the
  // one-block-at-a-time path): blocks are laid out contiguously, each starting at a multiple of 64 positions, in
  // gathered copies Es / Et of the resident matrices.
  std::vector<uint8_t> seen(D, 0);
  bool runs = !getenv("GPCA_DEBUG_NO_SHIFT_COPY");
  uint64_t max_m = 0, n_listed = 0;
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    if (m == 0) return fail(c, GPCA_ERR_INVALID, "empty LD block");
    max_m = std::max(max_m, m);
    for (uint64_t j = block_offsets[b]; j < block_offsets[b + 1]; ++j) {
      const uint64_t id = block_snp_ids[j];
      if (id >= D || seen[id]) return fail(c, GPCA_ERR_INVALID, "block SNP id out of range or listed twice");
      seen[id] = 1;
      if (j > block_offsets[b] && id != block_snp_ids[j - 1] + 1) runs = false;
    }
    n_listed += m;
  }
  const bool all_covered = n_listed == D;
  const uint32_t lp_cap = (uint32_t)std::min<uint64_t>(cpb_max + cfg->local_oversampling, std::min<uint64_t>(max_m, Ns));
  // All LD blocks in one launch per stage (integer engine, item mode) when the shapes allow it; otherwise one block
  // at a time through the generic sketch entry (any engine, missing calls, tiny inputs).
  const bool batched = c->batch_blocks && c->engine == 2 && !c->any_missing && sketch_i8_batch_supported(c) &&
                       lp_cap <= 32 && cpb_max <= 32 && Ns >= 128 && N >= 128 && D >= 128 && D / 4 < (1ull << 31) &&
                       n_blocks < (1ull << 24);
  const bool gs_whole = c->gs_win_rows == 0 || c->gs_res_rows >= D;
  const bool id_order = runs && all_covered && batched && (Ns != N || gs_whole) && !getenv("GPCA_DEBUG_NO_ID_ORDER");
  if (!id_order && !gs_whole)
    return fail(c, GPCA_ERR_OOM,
                "EigenSNP on a partly resident SNP-major matrix needs LD blocks that are runs of consecutive PCA SNPs "
                "covering every PCA SNP, without missing calls (otherwise: more GPUs or a smaller memory reserve)");
  std::vector<uint64_t> boff(n_blocks + 1, 0);     // first position of every block
  std::vector<int64_t> id_of_slot;                 // slot order only: PcaSnpId of a position (-1 = padding)
  uint64_t P = 0;                                  // number of positions
  if (id_order) {
    for (uint64_t b = 0; b < n_blocks; ++b) boff[b] = block_snp_ids[block_offsets[b]];
    boff[n_blocks] = D;
    P = D;
  } else {
    id_of_slot.reserve(D + 64 * n_blocks);
    for (uint64_t b = 0; b < n_blocks; ++b) {
      boff[b] = id_of_slot.size();
      for (uint64_t j = block_offsets[b]; j < block_offsets[b + 1]; ++j) id_of_slot.push_back((int64_t)block_snp_ids[j]);
      while (id_of_slot.size() % 64) id_of_slot.push_back(-1);
    }
    boff[n_blocks] = id_of_slot.size();
    P = id_of_slot.size();
  }
  std::vector<uint32_t> blk_of_pos(P, 0xffffffffu);      // block of a position (padding: none)
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    std::fill(blk_of_pos.begin() + boff[b], blk_of_pos.begin() + boff[b] + m, (uint32_t)b);
  }

  // ---- per-position vectors and the matrices over positions -----------------------------------------------------------
  PoolBuf<float> d_inv_slot(&c->es_pool), d_mu_slot(&c->es_pool);
  PoolBuf<int64_t> d_slot(&c->es_pool), d_sub(&c->es_pool);
  PoolBuf<uint32_t> d_blk_of_pos(&c->es_pool);
  GPCA_CUDA_TRY(c, d_sub.alloc(Ns));
  GPCA_CUDA_TRY(c, d_blk_of_pos.alloc(P));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_sub.p, sub.data(), Ns * 8, cudaMemcpyHostToDevice, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_blk_of_pos.p, blk_of_pos.data(), P * 4, cudaMemcpyHostToDevice, c->stream));
  std::vector<float> inv_s, mu_s;
  if (!id_order) {
    inv_s.assign(P, 0.f);
    mu_s.assign(P, 0.f);
    for (uint64_t q = 0; q < P; ++q)
      if (id_of_slot[q] >= 0) {
        const float sd = c->h_sd[id_of_slot[q]], mean = c->h_mean[id_of_slot[q]];
        if (!(std::fabs(sd) < 1e-9f)) {
          inv_s[q] = 1.0f / sd;
          mu_s[q] = mean * inv_s[q];
        }
      }
    GPCA_CUDA_TRY(c, d_inv_slot.alloc(P));
    GPCA_CUDA_TRY(c, d_mu_slot.alloc(P));
    GPCA_CUDA_TRY(c, d_slot.alloc(P));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_inv_slot.p, inv_s.data(), P * 4, cudaMemcpyHostToDevice, c->stream));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_mu_slot.p, mu_s.data(), P * 4, cudaMemcpyHostToDevice, c->stream));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_slot.p, id_of_slot.data(), P * 8, cudaMemcpyHostToDevice, c->stream));
  }
  const float* d_inv_p = id_order ? c->d_inv_sd.p : d_inv_slot.p;     // 1/sd, mean/sd per position
  const float* d_mu_p = id_order ? c->d_mu_inv_sd.p : d_mu_slot.p;
  struct { const float* p; } d_inv{d_inv_p}, d_mu{d_mu_p};

  // the gathered / subset copies are kept in the context between calls (allocating and freeing tens of GB per call
  // costs more than every kernel of a call together); gpca_load_* / gpca_ingest_bed release them
  DevBuf<uint8_t>&es_store = c->es_store, &et_store = c->et_store, &ets_store = c->ets_store, &ess_store = c->ess_store;
  PackedMat Es, Et, Ets, Ess;      // SNP-major / sample-major over positions, and their N_s-sample subsets
  if (id_order) {
    Es = c->Gs; Es.avail = c->Gs.pitch;      // (possibly only partly resident: read through for_each_gs_segment)
    Et = c->Gt; Et.avail = c->Gt.pitch;
    stage("  allocations + tables");
  } else {
    Es.rows = P; Es.cols = N; Es.pitch = c->Gs.pitch;
    Et.rows = N; Et.cols = P; Et.pitch = round_up((P + 3) / 4, 128);
    GPCA_CUDA_TRY(c, es_store.alloc(Es.pitch * Es.rows));
    GPCA_CUDA_TRY(c, et_store.alloc(Et.pitch * Et.rows));
    Es.p = es_store.p;
    Et.p = et_store.p;
    stage("  allocations + tables");
    GPCA_TRY(launch_gather_rows(c, c->Gs, d_slot.p, Es));
    stage("  gather slots");
    if (runs) {
      // blocks that are runs of consecutive PcaSnpIds: the sample-major copy is a shifted copy of Gt's rows
      const uint64_t n_chunks = P / 64;
      std::vector<int64_t> h_first(n_chunks, -1);
      std::vector<uint32_t> h_count(n_chunks, 0);
      for (uint64_t b = 0; b < n_blocks; ++b) {
        const uint64_t m = block_offsets[b + 1] - block_offsets[b];
        const uint64_t id0 = block_snp_ids[block_offsets[b]];
        for (uint64_t q = 0; q * 64 < m; ++q) {
          h_first[boff[b] / 64 + q] = (int64_t)(id0 + q * 64);
          h_count[boff[b] / 64 + q] = (uint32_t)std::min<uint64_t>(64, m - q * 64);
        }
      }
      PoolBuf<int64_t> d_first(&c->es_pool);
      PoolBuf<uint32_t> d_count(&c->es_pool);
      GPCA_CUDA_TRY(c, d_first.alloc(n_chunks));
      GPCA_CUDA_TRY(c, d_count.alloc(n_chunks));
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_first.p, h_first.data(), n_chunks * 8, cudaMemcpyHostToDevice, c->stream));
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_count.p, h_count.data(), n_chunks * 4, cudaMemcpyHostToDevice, c->stream));
      const uint64_t total = Et.rows * (Et.pitch / 16);
      const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->sm_count * 16);
      shift_fields_kernel<<<grid, 256, 0, c->stream>>>(c->Gt.p, c->Gt.pitch, d_first.p, d_count.p, Et.p, Et.pitch,
                                                       Et.rows, n_chunks);
      KCHECK(c);
      GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // (host tables above go out of scope)
    } else {
      GPCA_TRY(launch_transpose(c, Es, Et));
    }
    stage("  transpose");
  }
  if (Ns == N) {
    Ets = Et;
    Ess = Es;
  } else {
    Ets.rows = Ns; Ets.cols = P; Ets.pitch = Et.pitch; Ets.avail = Et.pitch;
    Ess.rows = P; Ess.cols = Ns; Ess.pitch = round_up((Ns + 3) / 4, 128); Ess.avail = Ess.pitch;
    GPCA_CUDA_TRY(c, ets_store.alloc(Ets.pitch * Ets.rows));
    GPCA_CUDA_TRY(c, ess_store.alloc(Ess.pitch * Ess.rows));
    Ets.p = ets_store.p;
    Ess.p = ess_store.p;
    // (id order: the subset copies depend only on the resident matrices and the subset -- a repeated call on the same
    //  data finds them in place)
    const bool cached = id_order && c->es_sub_copies_version == c->data_version && c->es_sub_copies_ns == Ns &&
                        c->es_sub_copies_seed == seed && !getenv("GPCA_DEBUG_NO_SUBSET_CACHE");
    if (!cached) {
      GPCA_TRY(launch_gather_rows(c, Et, d_sub.p, Ets));
      GPCA_TRY(launch_transpose(c, Ets, Ess));
      c->es_sub_copies_version = id_order ? c->data_version : 0;
      c->es_sub_copies_ns = Ns;
      c->es_sub_copies_seed = seed;
    }
  }
  const uint64_t Ds = P;      // (positions; the name the per-block code below uses)
  const std::vector<uint64_t>& off = boff;

  Small s;
  GPCA_TRY(get_small(c, s));
  stage("slot/subset copies");

  // ---- 2. local bases -------------------------------------------------------------------------------------
  const bool orth_both = getenv("GPCA_DEBUG_LOCAL_ORTH_BOTH") != nullptr;   // (A/B: re-orthonormalise both sides)
  std::vector<uint32_t> cp(n_blocks);
  std::vector<uint64_t> roff(n_blocks + 1, 0);
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    cp[b] = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(cpb_max, m), Ns);
    roff[b + 1] = roff[b] + cp[b];
  }
  const uint64_t R = roff[n_blocks];
  PoolBuf<float> Ubuf(&c->es_pool), Yb(&c->es_pool), Zb(&c->es_pool);
  GPCA_CUDA_TRY(c, Ubuf.alloc(Ds * cpb_max));
  GPCA_CUDA_TRY(c, cudaMemsetAsync(Ubuf.p, 0, Ds * cpb_max * sizeof(float), c->stream));
  std::vector<uint32_t> lpv(n_blocks);
  uint32_t lp_max = 0;
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    lpv[b] = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(cp[b] + cfg->local_oversampling, m), Ns);
    lp_max = std::max(lp_max, lpv[b]);
  }
  const uint64_t rgN = (N + 255) / 256, rgS = (Ns + 255) / 256;
  PoolBuf<SketchBatchBlock> d_blkY(&c->es_pool);     // operands indexed by the block's SNPs (K = block SNPs)
  std::vector<SketchBatchBlock> blkY;
  uint32_t img_stages_Y = 0;
  if (batched) {
    blkY.resize(n_blocks);
    for (uint64_t b = 0; b < n_blocks; ++b) {
      const uint64_t m = block_offsets[b + 1] - block_offsets[b];
      // the K range starts at the 64-field (16-byte) boundary at or below the block's first position; the operand rows
      // in front of the block are zeroed in the image (slot order: blocks start on such a boundary, kskip = 0)
      const uint64_t ka = off[b] & ~63ull;
      blkY[b].fe_off = ka;
      blkY[b].kskip = (uint32_t)(off[b] - ka);
      blkY[b].K = (uint32_t)(off[b] - ka + m);
      blkY[b].nst = (blkY[b].K + 255) / 256;
      blkY[b].img_st0 = img_stages_Y;
      img_stages_Y += blkY[b].nst;
    }
  }
  if (batched) {
    const uint32_t LD = lp_max;
    const uint32_t nstS = (uint32_t)((Ns + 255) / 256);
    if ((uint64_t)n_blocks * nstS > 0x7fffffffull || n_blocks * rgS > 0x7fffffffull)
      return fail(c, GPCA_ERR_INVALID, "too many LD blocks for one batch");
    std::vector<DenseProb> yprob(n_blocks), zprob(n_blocks);
    std::vector<uint32_t> streams(n_blocks);
    std::vector<uint64_t> uoffs(n_blocks);
    std::vector<SketchBatchBlock> blkZ(n_blocks);
    std::vector<I8Item> it1, it2;
    for (uint64_t b = 0; b < n_blocks; ++b) {
      const uint64_t m = block_offsets[b + 1] - block_offsets[b];
      yprob[b] = {off[b] * LD, (uint32_t)m, lpv[b]};
      zprob[b] = {b * Ns * LD, (uint32_t)Ns, lpv[b]};
      streams[b] = (uint32_t)(STREAM_LOCAL0 + c->shard_offset + block_snp_ids[block_offsets[b]]);
      uoffs[b] = off[b] * cpb_max;
      blkY[b].bin_off = (off[b] & ~63ull) * LD;
      blkY[b].l = lpv[b];
      blkZ[b].bin_off = b * Ns * LD;
      blkZ[b].kskip = 0;
      blkZ[b].fe_off = 0;
      blkZ[b].K = (uint32_t)Ns;
      blkZ[b].l = lpv[b];
      blkZ[b].img_st0 = (uint32_t)(b * nstS);
      blkZ[b].nst = nstS;
      for (uint64_t r0 = 0; r0 < m; r0 += 256) {   // rows = the block's SNPs, K = subset samples
        I8Item it;
        it.row0 = (uint32_t)(off[b] + r0);
        it.nrows_l = (uint32_t)std::min<uint64_t>(256, m - r0) | (lpv[b] << 16);
        it.kbyte0 = 0;
        it.nst = nstS;
        it.img_st0 = blkZ[b].img_st0;
        it.blk = (uint32_t)b;
        const uint64_t oo = (off[b] + r0) * LD;
        it.out_off_lo = (uint32_t)oo;
        it.out_off_hi = (uint32_t)(oo >> 32);
        it1.push_back(it);
      }
    }
    it2.reserve(rgS * n_blocks);
    for (uint64_t rg = 0; rg < rgS; ++rg)          // rows = subset samples, K = the block's SNPs
      for (uint64_t b = 0; b < n_blocks; ++b) {
        I8Item it;
        it.row0 = (uint32_t)(rg * 256);
        it.nrows_l = (uint32_t)std::min<uint64_t>(256, Ns - rg * 256) | (lpv[b] << 16);
        it.kbyte0 = (uint32_t)((off[b] & ~63ull) / 4);
        it.nst = blkY[b].nst;
        it.img_st0 = blkY[b].img_st0;
        it.blk = (uint32_t)b;
        const uint64_t oo = (b * Ns + rg * 256) * LD;
        it.out_off_lo = (uint32_t)oo;
        it.out_off_hi = (uint32_t)(oo >> 32);
        it2.push_back(it);
      }
    PoolBuf<DenseProb> d_yprob(&c->es_pool), d_zprob(&c->es_pool);
    PoolBuf<uint32_t> d_streams(&c->es_pool), d_cp(&c->es_pool);
    PoolBuf<uint64_t> d_uoffs(&c->es_pool);
    PoolBuf<SketchBatchBlock> d_blkZ(&c->es_pool);
    PoolBuf<I8Item> d_it1(&c->es_pool), d_it2(&c->es_pool);
    PoolBuf<float> Yall(&c->es_pool), Zall(&c->es_pool);
    GPCA_CUDA_TRY(c, d_yprob.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_zprob.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_streams.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_cp.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_uoffs.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_blkZ.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_blkY.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_it1.alloc(it1.size()));
    GPCA_CUDA_TRY(c, d_it2.alloc(it2.size()));
    GPCA_CUDA_TRY(c, Yall.alloc(Ds * LD));
    GPCA_CUDA_TRY(c, Zall.alloc(n_blocks * Ns * LD));
    auto up = [&](void* dst, const void* src, size_t bytes) {
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    GPCA_CUDA_TRY(c, up(d_yprob.p, yprob.data(), n_blocks * sizeof(DenseProb)));
    GPCA_CUDA_TRY(c, up(d_zprob.p, zprob.data(), n_blocks * sizeof(DenseProb)));
    GPCA_CUDA_TRY(c, up(d_streams.p, streams.data(), n_blocks * sizeof(uint32_t)));
    GPCA_CUDA_TRY(c, up(d_cp.p, cp.data(), n_blocks * sizeof(uint32_t)));
    GPCA_CUDA_TRY(c, up(d_uoffs.p, uoffs.data(), n_blocks * sizeof(uint64_t)));
    GPCA_CUDA_TRY(c, up(d_blkZ.p, blkZ.data(), n_blocks * sizeof(SketchBatchBlock)));
    GPCA_CUDA_TRY(c, up(d_blkY.p, blkY.data(), n_blocks * sizeof(SketchBatchBlock)));
    GPCA_CUDA_TRY(c, up(d_it1.p, it1.data(), it1.size() * sizeof(I8Item)));
    GPCA_CUDA_TRY(c, up(d_it2.p, it2.data(), it2.size() * sizeof(I8Item)));
    GPCA_CUDA_TRY(c, cudaMemsetAsync(Yall.p, 0, Ds * LD * sizeof(float), c->stream));   // pad slots stay zero
    DenseBatchWs ws;
    GPCA_TRY(dense_batch_ws(c, (uint32_t)n_blocks, ws));
    double bytes_blocks = 0.0;
    for (uint64_t b = 0; b < n_blocks; ++b) bytes_blocks += (double)(block_offsets[b + 1] - block_offsets[b]);
    SketchBatch p1;   // Y_b = X_b Z_b
    p1.G = Ess; p1.G.avail = Ess.pitch;
    p1.d_items = d_it1.p; p1.n_items = (uint32_t)it1.size();
    p1.d_blocks = d_blkZ.p; p1.n_blocks = (uint32_t)n_blocks;
    p1.total_img_stages = (uint32_t)(n_blocks * nstS); p1.max_K = (uint32_t)Ns;
    p1.Bin = Zall.p; p1.ld = LD; p1.f = nullptr; p1.e = nullptr; p1.a = d_inv.p; p1.b = d_mu.p;
    p1.out = Yall.p; p1.ldo = LD;
    p1.bytes = bytes_blocks * (double)((Ns + 3) / 4);
    SketchBatch p2;   // Z_b = X_b^T Y_b
    p2.G = Ets; p2.G.avail = Ets.pitch;
    p2.d_items = d_it2.p; p2.n_items = (uint32_t)it2.size();
    p2.d_blocks = d_blkY.p; p2.n_blocks = (uint32_t)n_blocks;
    p2.total_img_stages = img_stages_Y; p2.max_K = (uint32_t)max_m + 64;
    p2.Bin = Yall.p; p2.ld = LD; p2.f = d_inv.p; p2.e = d_mu.p; p2.a = nullptr; p2.b = nullptr;
    p2.out = Zall.p; p2.ldo = LD;
    p2.bytes = (double)Ns * bytes_blocks / 4.0;
    const uint32_t nb = (uint32_t)n_blocks;
    GPCA_TRY(launch_gaussian_batch(c, Zall.p, LD, d_zprob.p, nb, Ns, seed, d_streams.p));
    GPCA_TRY(timed_sketch_batch(c, p1));                                               // Y = X Omega
    for (uint32_t it = 0; it < cfg->local_power_iters; ++it) {
      GPCA_TRY(orthonormalize_batch(c, Yall.p, LD, d_yprob.p, nb, max_m, lp_max, ws));
      GPCA_TRY(timed_sketch_batch(c, p2));                                             // Z = X^T Q
      // (range(X Z) does not depend on a column transform of Z: as in gpca_rfit only one side -- the small one, the
      //  block's SNPs -- is re-orthonormalised per iteration; CholeskyQR2 on the N_s-row iterates of all blocks was
      //  four Gram and four Y.T launches over 8 M rows, half of this stage)
      if (orth_both) GPCA_TRY(orthonormalize_batch(c, Zall.p, LD, d_zprob.p, nb, Ns, lp_max, ws));
      GPCA_TRY(timed_sketch_batch(c, p1));                                             // Y = X Z
    }
    GPCA_TRY(orthonormalize_batch(c, Yall.p, LD, d_yprob.p, nb, max_m, lp_max, ws));
    GPCA_TRY(timed_sketch_batch(c, p2));                                               // B^T = X^T Q   [Ns x lp]
    int np = 1;
    GPCA_TRY(launch_gram_batch(c, Zall.p, LD, d_zprob.p, nb, Ns, &np));
    GPCA_TRY(launch_chol_orth_batch(c, np, d_zprob.p, nb, 0.0, ws));                   // (sums the partials -> ws.G)
    GPCA_TRY(launch_jacobi_eigh_batch(c, d_zprob.p, nb, ws, false));
    GPCA_TRY(launch_rotation_batch(c, d_zprob.p, nb, d_cp.p, ws));
    GPCA_TRY(launch_apply_right_batch(c, Yall.p, LD, d_yprob.p, nb, max_m, ws.T, d_cp.p, cpb_max, Ubuf.p, d_uoffs.p,
                                      cpb_max));                                       // U_p = Q U_b[:, :c_p]
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // host tables and the temporaries above go out of scope
  } else {
  GPCA_CUDA_TRY(c, Yb.alloc(max_m * 64));
  GPCA_CUDA_TRY(c, Zb.alloc(Ns * 64));
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    const uint64_t o = off[b];
    const uint32_t lp = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(cp[b] + cfg->local_oversampling, m), Ns);
    SketchProblem p1;   // rows = block SNPs, K = subset samples
    p1.G.p = Ess.p + o * Ess.pitch; p1.G.pitch = Ess.pitch; p1.G.rows = m; p1.G.cols = Ns; p1.G.avail = Ess.pitch;
    p1.l = lp; p1.ld = lp; p1.f = nullptr; p1.e = nullptr; p1.a = d_inv.p + o; p1.b = d_mu.p + o; p1.ldo = lp;
    SketchProblem p2;   // rows = subset samples, K = block SNPs
    p2.G.p = Ets.p + o / 4; p2.G.pitch = Ets.pitch; p2.G.rows = Ns; p2.G.cols = m; p2.G.avail = Ets.pitch - o / 4;
    p2.l = lp; p2.ld = lp; p2.f = d_inv.p + o; p2.e = d_mu.p + o; p2.a = nullptr; p2.b = nullptr; p2.ldo = lp;
    const uint64_t first_id = c->shard_offset + block_snp_ids[block_offsets[b]];
    GPCA_TRY(launch_gaussian(c, Zb.p, Ns, lp, lp, seed, (uint32_t)(STREAM_LOCAL0 + first_id), 0));
    p1.Bin = Zb.p; p1.out = Yb.p;
    GPCA_TRY(timed_sketch(c, p1));                               // Y = X Omega
    for (uint32_t it = 0; it < cfg->local_power_iters; ++it) {
      GPCA_TRY(orthonormalize(c, Yb.p, m, lp, lp, false, s));
      p2.Bin = Yb.p; p2.out = Zb.p;
      GPCA_TRY(timed_sketch(c, p2));                             // Z = X^T Q
      if (orth_both) GPCA_TRY(orthonormalize(c, Zb.p, Ns, lp, lp, false, s));
      GPCA_TRY(timed_sketch(c, p1));                             // Y = X Z
    }
    GPCA_TRY(orthonormalize(c, Yb.p, m, lp, lp, false, s));
    p2.Bin = Yb.p; p2.out = Zb.p;
    GPCA_TRY(timed_sketch(c, p2));                               // B^T = X^T Q   [Ns x lp]
    GPCA_TRY(launch_gram(c, Zb.p, Ns, lp, lp, s.G));
    GPCA_TRY(launch_jacobi_eigh(c, s.G, lp, s.evals, s.evecs));
    GPCA_TRY(launch_rotation_transform(c, s.evals, s.evecs, lp, cp[b], s.T, false));
    GPCA_TRY(launch_apply_right(c, Yb.p, m, lp, lp, s.T, cp[b], Ubuf.p + o * cpb_max, cpb_max));   // U_p = Q U_b[:, :c_p]
  }
  }

  stage("local bases");
  // ---- 3. condensed features (all N samples), Cn [N x R], then column standardisation -----------------------
  if (R == 0) return fail(c, GPCA_ERR_INVALID, "no condensed features");
  // (row stride a multiple of 4 floats: 16-byte row pitch for the TMA loads of the dense products; pad columns zero)
  const uint32_t ldc = (uint32_t)round_up(R, 4);
  DevBuf<float>& Cn = c->es_cn;
  GPCA_CUDA_TRY(c, Cn.alloc(N * ldc));
  if (ldc != R)
    GPCA_CUDA_TRY(c, cudaMemset2DAsync(Cn.p + R, (size_t)ldc * 4, 0, (size_t)(ldc - R) * 4, N, c->stream));
  if (batched) {
    // Consecutive blocks share a work item while their component counts fit the 32 accumulator columns: the operand
    // of a group is block-diagonal (rows = the group's slot range, block p's components in its own columns), so one
    // pass over the group's K range leaves every block's features in adjacent columns of Cn.  With one item per
    // (row group, block) an item was two stages long and the launch was bound by the per-item hand-overs
    // (13.6 ms for 10.9 GB at 500,000 x 87,500 / 212 blocks); groups of 4 blocks (4 x 7 columns) run 7-stage items.
    struct Grp {
      uint64_t b0, b1;
      uint32_t l;
    };
    std::vector<Grp> grps;
    const bool grouping = !getenv("GPCA_DEBUG_NO_GROUPS");
    for (uint64_t b = 0; b < n_blocks; ++b) {
      // (id order: a group's K range is one run of positions, so only blocks that follow each other there can share it;
      //  slot order lays consecutive blocks out next to each other, separated by zero padding)
      const bool adjacent = b > 0 && (!id_order || off[b] == off[b - 1] + (block_offsets[b] - block_offsets[b - 1]));
      if (grouping && !grps.empty() && adjacent && grps.back().l + cp[b] <= 32) {
        grps.back().b1 = b + 1;
        grps.back().l += cp[b];
      } else {
        grps.push_back({b, b + 1, cp[b]});
      }
    }
    const uint64_t n_grps = grps.size();
    if (rgN * n_grps > 0x7fffffffull) return fail(c, GPCA_ERR_INVALID, "too many condensed-feature work items");
    // block-diagonal operand [positions x 32]: a position of block p holds U_p's row in columns [col0(p), col0(p) + c_p)
    std::vector<uint32_t> h_col0(n_blocks, 0);
    std::vector<SketchBatchBlock> blkG(n_grps);
    std::vector<uint64_t> grp_ka(n_grps);
    uint32_t img_stages_G = 0, max_KG = 0;
    for (uint64_t g = 0; g < n_grps; ++g) {
      uint32_t col = 0;
      for (uint64_t b = grps[g].b0; b < grps[g].b1; ++b) {
        h_col0[b] = col;
        col += cp[b];
      }
      const uint64_t bl = grps[g].b1 - 1;
      const uint64_t ka = off[grps[g].b0] & ~63ull;      // K range from the 64-field boundary below the group
      const uint64_t Kg = off[bl] + (block_offsets[bl + 1] - block_offsets[bl]) - ka;
      grp_ka[g] = ka;
      blkG[g].bin_off = ka * 32;
      blkG[g].fe_off = ka;
      blkG[g].kskip = (uint32_t)(off[grps[g].b0] - ka);
      blkG[g].K = (uint32_t)Kg;
      blkG[g].l = grps[g].l;
      blkG[g].nst = (uint32_t)((Kg + 255) / 256);
      blkG[g].img_st0 = img_stages_G;
      img_stages_G += blkG[g].nst;
      max_KG = std::max(max_KG, blkG[g].K);
    }
    PoolBuf<float> Ugrp(&c->es_pool);
    PoolBuf<uint32_t> d_col0(&c->es_pool), d_cpn(&c->es_pool);
    PoolBuf<SketchBatchBlock> d_blkG(&c->es_pool);
    GPCA_CUDA_TRY(c, Ugrp.alloc(Ds * 32));
    GPCA_CUDA_TRY(c, d_col0.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_cpn.alloc(n_blocks));
    GPCA_CUDA_TRY(c, d_blkG.alloc(n_grps));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_col0.p, h_col0.data(), n_blocks * 4, cudaMemcpyHostToDevice, c->stream));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_cpn.p, cp.data(), n_blocks * 4, cudaMemcpyHostToDevice, c->stream));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_blkG.p, blkG.data(), n_grps * sizeof(SketchBatchBlock), cudaMemcpyHostToDevice,
                                     c->stream));
    GPCA_CUDA_TRY(c, cudaMemsetAsync(Ugrp.p, 0, Ds * 32 * sizeof(float), c->stream));
    {
      const uint64_t tot = Ds * cpb_max;
      const int grid = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)c->sm_count * 8);
      block_diag_operand_kernel<<<grid, 256, 0, c->stream>>>(Ubuf.p, cpb_max, d_blk_of_pos.p, d_col0.p, d_cpn.p, Ds, Ugrp.p);
      KCHECK(c);
    }
    std::vector<I8Item> itc;
    itc.reserve(rgN * n_grps);
    for (uint64_t rg = 0; rg < rgN; ++rg)
      for (uint64_t g = 0; g < n_grps; ++g) {
        I8Item it;
        it.row0 = (uint32_t)(rg * 256);
        it.nrows_l = (uint32_t)std::min<uint64_t>(256, N - rg * 256) | (blkG[g].l << 16);
        it.kbyte0 = (uint32_t)(grp_ka[g] / 4);
        it.nst = blkG[g].nst;
        it.img_st0 = blkG[g].img_st0;
        it.blk = (uint32_t)g;
        const uint64_t oo = rg * 256 * (uint64_t)ldc + roff[grps[g].b0];
        it.out_off_lo = (uint32_t)oo;
        it.out_off_hi = (uint32_t)(oo >> 32);
        itc.push_back(it);
      }
    PoolBuf<I8Item> d_itc(&c->es_pool);
    GPCA_CUDA_TRY(c, d_itc.alloc(itc.size()));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_itc.p, itc.data(), itc.size() * sizeof(I8Item), cudaMemcpyHostToDevice, c->stream));
    SketchBatch pc;   // C_b = X_b^T U_b for all N samples, written into the block's columns of Cn
    pc.G = Et; pc.G.avail = Et.pitch;
    pc.d_items = d_itc.p; pc.n_items = (uint32_t)itc.size();
    pc.d_blocks = d_blkG.p; pc.n_blocks = (uint32_t)n_grps;
    pc.total_img_stages = img_stages_G; pc.max_K = max_KG;
    pc.Bin = Ugrp.p; pc.ld = 32; pc.f = d_inv.p; pc.e = d_mu.p; pc.a = nullptr; pc.b = nullptr;
    pc.out = Cn.p; pc.ldo = ldc;
    pc.bytes = (double)N * (double)D / 4.0;
    pc.wide_boxes = Et.pitch >= 32768 && max_KG >= 4 * 256;      // long items, row pitch of tens of KB: TMA-stream-bound
    GPCA_TRY(timed_sketch_batch(c, pc));
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  } else {
  for (uint64_t b = 0; b < n_blocks; ++b) {
    const uint64_t m = block_offsets[b + 1] - block_offsets[b];
    const uint64_t o = off[b];
    SketchProblem p2;
    p2.G.p = Et.p + o / 4; p2.G.pitch = Et.pitch; p2.G.rows = N; p2.G.cols = m; p2.G.avail = Et.pitch - o / 4;
    p2.l = cp[b]; p2.ld = cpb_max; p2.f = d_inv.p + o; p2.e = d_mu.p + o; p2.a = nullptr; p2.b = nullptr;
    p2.Bin = Ubuf.p + o * cpb_max; p2.out = Cn.p + roff[b]; p2.ldo = ldc;
    GPCA_TRY(timed_sketch(c, p2));
  }
  }
  stage("  condensed features");
  // column moments of the condensed matrix (one read-only sweep); its standardised form is never written
  PoolBuf<float> cmean(&c->es_pool), cinv(&c->es_pool), cmuinv(&c->es_pool);
  {
    GPCA_CUDA_TRY(c, cmean.alloc(R));
    GPCA_CUDA_TRY(c, cinv.alloc(R));
    GPCA_CUDA_TRY(c, cmuinv.alloc(R));
    int nparts = (int)std::min<uint64_t>((N + 2047) / 2048, 64);
    if (nparts < 1) nparts = 1;
    const uint64_t rpc = (N + nparts - 1) / nparts;
    nparts = (int)((N + rpc - 1) / rpc);
    PoolBuf<double> part(&c->es_pool);
    GPCA_CUDA_TRY(c, part.alloc((size_t)nparts * R * 2));
    dim3 g1(nparts, (unsigned)((R + 255) / 256));
    col_moments_kernel<<<g1, 256, 0, c->stream>>>(Cn.p, N, (uint32_t)R, ldc, rpc, part.p);
    KCHECK(c);
    col_finalize_kernel<<<(unsigned)((R + 255) / 256), 256, 0, c->stream>>>(part.p, nparts, N, (uint32_t)R, cmean.p, cinv.p,
                                                                           cmuinv.p);
    KCHECK(c);
  }

  stage("condense + standardise");
  // ---- 4. global randomized SVD of the condensed matrix (rows of C^T sharded by rank) ------------------------
  uint64_t R_total = R;
  if (c->sharded()) {   // total condensed rows over all shards (f64 scalar through the exchange)
    PoolBuf<double> tmp(&c->es_pool);
    GPCA_CUDA_TRY(c, tmp.alloc(1));
    const double rr = (double)R;
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(tmp.p, &rr, 8, cudaMemcpyHostToDevice, c->stream));
    GPCA_TRY(driver_allreduce(c, tmp.p, 1, 1));
    double rt = 0;
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(&rt, tmp.p, 8, cudaMemcpyDeviceToHost, c->stream));
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    R_total = (uint64_t)(rt + 0.5);
  }
  const uint32_t lg = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(k_req + cfg->global_oversampling, R_total), N);
  const uint32_t k = std::min<uint32_t>(k_req, lg);
  PoolBuf<float> Om(&c->es_pool), Yg(&c->es_pool), Zg(&c->es_pool), V(&c->es_pool), L(&c->es_pool), Sc(&c->es_pool);
  GPCA_CUDA_TRY(c, Om.alloc(R * lg));
  GPCA_CUDA_TRY(c, Yg.alloc(N * lg));
  GPCA_CUDA_TRY(c, Zg.alloc(R * lg));
  {
    std::vector<uint64_t> keys(R);
    for (uint64_t b = 0; b < n_blocks; ++b)
      for (uint32_t j = 0; j < cp[b]; ++j)
        keys[roff[b] + j] = (c->shard_offset + block_snp_ids[block_offsets[b]]) * 64ull + j;
    PoolBuf<uint64_t> d_keys(&c->es_pool);
    GPCA_CUDA_TRY(c, d_keys.alloc(R));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_keys.p, keys.data(), R * 8, cudaMemcpyHostToDevice, c->stream));
    const uint64_t tot = R * lg;
    const int grid = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)c->sm_count * 16);
    gaussian_keyed_kernel<<<grid, 256, 0, c->stream>>>(Om.p, d_keys.p, R, lg, seed, STREAM_GLOBAL);
    KCHECK(c);
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  stage("  omega");
  // products with the standardised condensed matrix Cz = (Cn - 1 mean^T) diag(1/sd), through its scales:
  //   Cz   W = Cn (W / sd) - 1 (mean / sd)^T W             [N x lg]   (rows = samples)
  //   Cz^T Y = diag(1/sd) Cn^T Y - (mean / sd) (1^T Y)      [R x lg]   (rows = condensed features)
  auto cz_times = [&](const float* w, float* y) -> int {
    DenseProduct dp{Cn.p, N, R, ldc, false, w, lg, lg, cinv.p, cmuinv.p, nullptr, nullptr, y, lg};
    return launch_dense_product(c, dp);
  };
  auto czt_times = [&](const float* y, float* z) -> int {
    DenseProduct dp{Cn.p, N, R, ldc, true, y, lg, lg, nullptr, nullptr, cinv.p, cmuinv.p, z, lg};
    return launch_dense_product(c, dp);
  };
  GPCA_TRY(cz_times(Om.p, Yg.p));                                        // Yg = Cz Omega
  GPCA_TRY(driver_allreduce(c, Yg.p, N * lg, 0));
  stage("  first gemm");
  for (uint32_t it = 0; it < cfg->global_power_iters; ++it) {
    // (only the small side -- the R condensed rows -- is re-orthonormalised inside the iteration; the N-row iterate is
    //  orthonormalised once, in front of the projection)
    if (orth_both) GPCA_TRY(orthonormalize(c, Yg.p, N, lg, lg, false, s));
    GPCA_TRY(czt_times(Yg.p, Zg.p));                                     // Zg = Cz^T Yg
    GPCA_TRY(orthonormalize(c, Zg.p, R, lg, lg, true, s));
    GPCA_TRY(cz_times(Zg.p, Yg.p));                                      // Yg = Cz Q_z
    GPCA_TRY(driver_allreduce(c, Yg.p, N * lg, 0));
    stage("  power iteration");
  }
  GPCA_TRY(orthonormalize(c, Yg.p, N, lg, lg, false, s));
  GPCA_TRY(czt_times(Yg.p, Zg.p));                                       // B = Cz^T Q
  GPCA_TRY(launch_gram(c, Zg.p, R, lg, lg, s.G));
  GPCA_TRY(driver_allreduce(c, s.G, (uint64_t)lg * lg, 1));
  GPCA_TRY(launch_jacobi_eigh(c, s.G, lg, s.evals, s.evecs));
  GPCA_TRY(launch_rotation_transform(c, s.evals, s.evecs, lg, k, s.T, false));
  GPCA_CUDA_TRY(c, V.alloc(N * k));
  GPCA_TRY(launch_apply_right(c, Yg.p, N, lg, lg, s.T, k, V.p, k));     // V0 = Q V_b[:, :k]

  stage("global rSVD");
  // ---- 5. refinement on the full genotype matrix ---------------------------------------------------------------
  // rows of the loadings: slots, or -- id_order -- the PCA SNPs themselves (resident Gs / Gt and their 1/sd, mean/sd)
  const uint64_t Dl = id_order ? D : Ds;
  const float* l_inv = id_order ? c->d_inv_sd.p : d_inv.p;
  const float* l_mu = id_order ? c->d_mu_inv_sd.p : d_mu.p;
  GPCA_CUDA_TRY(c, L.alloc(Dl * k));
  GPCA_CUDA_TRY(c, Sc.alloc(N * k));
  SketchProblem f1;   // L = S V   (rows = slots / SNPs, K = all samples)
  f1.G = Es; f1.G.avail = Es.pitch;      // (id order: the resident matrix, read segment by segment below)
  f1.l = k; f1.ld = k; f1.f = nullptr; f1.e = nullptr; f1.a = l_inv; f1.b = l_mu; f1.ldo = k;
  SketchProblem f2;   // Sc = S^T L (rows = samples, K = slots / SNPs)
  if (id_order) {
    f2.G = c->Gt; f2.G.avail = c->Gt.pitch;
  } else {
    f2.G = Et; f2.G.avail = Et.pitch;
  }
  f2.l = k; f2.ld = k; f2.f = l_inv; f2.e = l_mu; f2.a = nullptr; f2.b = nullptr; f2.ldo = k;
  PoolBuf<double> d_lam(&c->es_pool);
  GPCA_CUDA_TRY(c, d_lam.alloc(64));
  const uint32_t passes = cfg->refine_pass_count;
  for (uint32_t pass = 0; pass < std::max<uint32_t>(passes, 1); ++pass) {
    f1.Bin = V.p; f1.out = L.p;
    if (id_order) {
      GPCA_TRY(for_each_gs_segment(c, [&](const PackedMat& g, uint64_t row0) -> int {
        SketchProblem q = f1;
        q.G = g;
        q.a = l_inv + row0;
        q.b = l_mu + row0;
        q.out = L.p + row0 * k;
        return timed_sketch(c, q);
      }));
    } else {
      GPCA_TRY(timed_sketch(c, f1));
    }
    if (passes == 0) {
      // no refinement requested: loadings = normalised S V0, singular values = column norms
      GPCA_TRY(launch_gram(c, L.p, Dl, k, k, s.G));
      GPCA_TRY(driver_allreduce(c, s.G, (uint64_t)k * k, 1));
      break;
    }
    GPCA_TRY(orthonormalize(c, L.p, Dl, k, k, true, s));
    f2.Bin = L.p; f2.out = Sc.p;
    GPCA_TRY(timed_sketch(c, f2));
    GPCA_TRY(driver_allreduce(c, Sc.p, N * (uint64_t)k, 0));
    GPCA_TRY(launch_gram(c, Sc.p, N, k, k, s.G));
    GPCA_TRY(launch_jacobi_eigh(c, s.G, k, s.evals, s.evecs));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_lam.p, s.evals, k * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    GPCA_TRY(launch_rotation_transform(c, s.evals, s.evecs, k, k, s.T, true));
    GPCA_TRY(launch_apply_right(c, Sc.p, N, k, k, s.T, k, V.p, k));          // V = Sc W / sigma (orthonormal)
    GPCA_TRY(launch_rotation_transform(c, s.evals, s.evecs, k, k, s.T, false));
    GPCA_TRY(launch_apply_right(c, L.p, Dl, k, k, s.T, k, L.p, k));          // loadings = L W
  }
  stage("refinement");
  std::vector<double> h_lam(k, 0.0);
  if (passes == 0) {
    std::vector<double> hg((size_t)k * k);
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(hg.data(), s.G, (size_t)k * k * 8, cudaMemcpyDeviceToHost, c->stream));
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    std::vector<double> t((size_t)k * k, 0.0);
    for (uint32_t j = 0; j < k; ++j) {
      h_lam[j] = hg[(size_t)j * k + j];
      t[(size_t)j * k + j] = h_lam[j] > 0 ? 1.0 / std::sqrt(h_lam[j]) : 0.0;
    }
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(s.T, t.data(), (size_t)k * k * 8, cudaMemcpyHostToDevice, c->stream));
    GPCA_TRY(launch_apply_right(c, L.p, Dl, k, k, s.T, k, L.p, k));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_lam.p, h_lam.data(), k * 8, cudaMemcpyHostToDevice, c->stream));
  } else {
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(h_lam.data(), d_lam.p, k * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  // scores = V * sigma
  {
    const uint64_t tot = N * (uint64_t)k;
    const int grid = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)c->sm_count * 8);
    scale_cols_sqrt_kernel<<<grid, 256, 0, c->stream>>>(V.p, N, k, d_lam.p);
    KCHECK(c);
  }
  // sign convention and the slot -> PcaSnpId scatter of the loadings happen on the device; the host receives the
  // final buffers only
  PoolBuf<int> d_flags(&c->es_pool);
  GPCA_CUDA_TRY(c, d_flags.alloc(64));
  GPCA_TRY(launch_sign_flags(c, V.p, N, k, k, d_flags.p));
  if (scores) {
    GPCA_TRY(launch_apply_flags(c, V.p, N, k, k, d_flags.p, V.p, nullptr));
    GPCA_TRY(download_results(c, V.p, N * k, scores, nullptr));
  }
  PoolBuf<float> Lout(&c->es_pool);
  if (loadings && id_order) {      // already in PcaSnpId order: only the sign flips are left
    GPCA_TRY(launch_apply_flags(c, L.p, D, k, k, d_flags.p, L.p, nullptr));
    GPCA_TRY(download_results(c, L.p, D * k, loadings, nullptr));
  } else if (loadings) {
    GPCA_CUDA_TRY(c, Lout.alloc(D * k));
    GPCA_CUDA_TRY(c, cudaMemsetAsync(Lout.p, 0, D * k * sizeof(float), c->stream));
    const uint64_t tot = Ds * (uint64_t)k;
    const int grid = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)c->sm_count * 8);
    scatter_loadings_kernel<<<grid, 256, 0, c->stream>>>(L.p, d_slot.p, Ds, k, d_flags.p, Lout.p);
    KCHECK(c);
    GPCA_TRY(download_results(c, Lout.p, D * k, loadings, nullptr));
  }
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (eigenvalues)
    for (uint32_t j = 0; j < k; ++j) eigenvalues[j] = h_lam[j] / (double)(N - 1);
  stage("outputs");
  if (diag) {
    char buf[512];
    std::string j = "{\n  \"producer\": \"";
    j += gpca_version();
    j += "\",\n";
    snprintf(buf, sizeof buf,
             "  \"num_qc_samples\": %llu,\n  \"num_pca_snps\": %llu,\n  \"num_ld_blocks\": %llu,\n"
             "  \"subset_samples\": %llu,\n  \"condensed_features\": %llu,\n  \"condensed_features_all_shards\": %llu,\n"
             "  \"global_sketch_columns\": %u,\n  \"components\": %u,\n",
             (unsigned long long)N, (unsigned long long)D, (unsigned long long)n_blocks, (unsigned long long)Ns,
             (unsigned long long)R, (unsigned long long)R_total, lg, k);
    j += buf;
    snprintf(buf, sizeof buf,
             "  \"layout\": \"%s\",\n  \"blocks_per_launch\": %s,\n  \"snp_major_rows_resident\": %llu,\n"
             "  \"shards\": %d,\n  \"shard_offset\": %llu,\n  \"refine_passes\": %u,\n",
             id_order ? "id order (resident matrices)" : "slot order (gathered copies)", batched ? "true" : "false",
             (unsigned long long)(c->gs_win_rows ? c->gs_res_rows : D), c->comm_world, (unsigned long long)c->shard_offset, passes);
    j += buf;
    snprintf(buf, sizeof buf,
             "  \"kernel_launches\": %llu,\n  \"collectives\": %llu,\n  \"sketch_passes\": %llu,\n"
             "  \"packed_genotype_bytes_streamed\": %.0f,\n  \"total_ms\": %.3f,\n",
             (unsigned long long)(c->launches - launches0), (unsigned long long)(c->collectives - collectives0),
             (unsigned long long)(c->sk_passes - sk_passes0), c->sk_bytes - sk_bytes0,
             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
    j += buf;
    j += "  \"eigenvalues\": [";
    for (uint32_t q = 0; q < k; ++q) {
      snprintf(buf, sizeof buf, "%s%.9g", q ? ", " : "", h_lam[q] / (double)(N - 1));
      j += buf;
    }
    j += "],\n  \"stages_ms\": [\n";
    for (size_t q = 0; q < diag_stages.size(); ++q) {
      snprintf(buf, sizeof buf, "    {\"stage\": \"%s\", \"ms\": %.3f}%s\n", diag_stages[q].first.c_str(), diag_stages[q].second,
               q + 1 < diag_stages.size() ? "," : "");
      j += buf;
    }
    j += "  ]\n}\n";
    c->es_diag_json = j;
  }
  if (k_out) *k_out = k;
  return GPCA_OK;
}

extern "C" const char* gpca_eigensnp_diagnostics(const gpca_ctx* c) { return c ? c->es_diag_json.c_str() : ""; }
