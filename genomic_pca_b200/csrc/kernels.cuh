// kernels.cuh -- host-callable launch wrappers (one per kernel family).
#pragma once
#include "common.cuh"

// ---- ingest (kernels_ingest.cu) ---------------------------------------------------------
// PLINK rows with arbitrary byte pitch -> 16B-aligned pitch, optional sample gather; pad fields = 01.
int launch_repitch_gather(gpca_ctx* c, const uint8_t* d_in, size_t in_pitch, uint64_t n_in_samples,
                          const int64_t* d_keep, uint64_t N, uint64_t M, uint8_t* d_out, size_t out_pitch);
// variant-major u8 dosages -> PLINK-coded rows (pad fields = 01)
int launch_u8_to_plink(gpca_ctx* c, const uint8_t* d_in, uint64_t N, uint64_t M, uint8_t* d_out, size_t out_pitch);
// K-a: per-row code counts over ALL fields of the pitch: out[row] = {n(01), n(10), n(11), 0}
int launch_bed_counts(gpca_ctx* c, const uint8_t* d_raw, size_t pitch, uint64_t M, uint4* d_out);
// gather rows + recode PLINK -> dosage-coded, zero pads: Gs
int launch_build_gs(gpca_ctx* c, const uint8_t* d_raw, size_t raw_pitch, const uint64_t* d_idx, PackedMat gs);
// streaming ingest: counts / recode straight from a staged chunk whose rows sit at the file's (unaligned) pitch
int launch_chunk_counts(gpca_ctx* c, const uint8_t* d_src, size_t pitch, uint64_t N, uint64_t M, uint4* out);
int launch_fetch_kept(gpca_ctx* c, const uint64_t* u_idx, const float* u_mean, const float* u_sd, const float* u_inv,
                      const float* u_mu, uint64_t kept, uint64_t* d_idx, float* d_mean, float* d_sd, float* d_inv,
                      float* d_mu);
int launch_build_gs_chunk(gpca_ctx* c, const uint8_t* d_src, size_t pitch, uint64_t N, uint64_t row_base,
                          const uint64_t* d_idx, uint64_t kept, uint8_t* gs, size_t gs_pitch, uint64_t dst_row0,
                          uint64_t res_rows, uint64_t win_rows);
// dst row i = src row d_idx[i] (idx < 0 -> zero row)
int launch_gather_rows(gpca_ctx* c, PackedMat src, const int64_t* d_idx, PackedMat dst);
// 2-bit transpose: Gt[n][d] = Gs[d][n]
int launch_transpose(gpca_ctx* c, PackedMat gs, PackedMat gt);
// standardized block (accessor parity): out[i][j] = fma(x, 1/sd, -mean/sd); flag set if any missing
int launch_std_block(gpca_ctx* c, PackedMat gs, uint64_t gs_rows, PackedMat gt, const float* d_mean, const float* d_sd, const uint64_t* d_ids,
                     uint64_t n_ids, const uint64_t* d_samp, uint64_t n_samp, float* d_out, int* d_missing_flag);

// synthetic Balding-Nichols genotypes in .bed layout, keyed by (seed, global SNP index, sample): benchmark input
int launch_synth_bed(gpca_ctx* c, uint8_t* d_out, uint64_t n_samples, uint64_t n_snps, uint64_t snp_offset, uint64_t seed,
                     uint32_t n_pops, float fst, float missing_rate, float fst_grade);

// ---- dense helpers (kernels_dense.cu) ---------------------------------------------------
// Gaussian test matrix: out[r][c] = N(0,1) keyed by (seed, stream, row0 + r, c); ld in floats
int launch_gaussian(gpca_ctx* c, float* d_out, uint64_t rows, uint32_t cols, uint32_t ld, uint64_t seed,
                    uint32_t stream, uint64_t row0);
// same matrix, and the operand statistics (column sums weighted by e, max |f o out|) left in the context for the
// sketch pass that consumes it (cols <= 32)
int launch_gaussian_with_stats(gpca_ctx* c, float* d_out, uint64_t rows, uint32_t cols, uint32_t ld, uint64_t seed,
                               uint32_t stream, uint64_t row0, const float* d_f, const float* d_e);
// context-held operand statistics (layout: [STATS_MAX_PARTS][32] f64 partial column sums, then the max-abs word)
constexpr int STATS_MAX_PARTS = 4096;
int stats_buffer(gpca_ctx* c, double** cpart, unsigned int** amax);
int stats_begin_produce(gpca_ctx* c, unsigned int* amax);
// G[l x l] (f64, row-major) = Y^T Y, Y [n x l] fp32 with row stride ld.  Deterministic two-stage.
int launch_gram(gpca_ctx* c, const float* d_y, uint64_t n, uint32_t l, uint32_t ld, double* d_g);
// G[l x l] (f64) = A^T B for two [n x l] fp32 matrices with the same row stride
int launch_cross_gram(gpca_ctx* c, const float* d_a, const float* d_b, uint64_t n, uint32_t l, uint32_t ld,
                      double* d_g);
// Y <- Y * T (T [l x l2] f64 row-major on device), in place allowed when d_out == d_y. out ld = ldo
int launch_apply_right(gpca_ctx* c, const float* d_y, uint64_t n, uint32_t l, uint32_t ld, const double* d_t,
                       uint32_t l2, float* d_out, uint32_t ldo);
// single-CTA Jacobi eigensolver on a symmetric l x l f64 matrix (l <= 64):
// evals descending, evecs as columns (row-major [l x l]).
int launch_jacobi_eigh(gpca_ctx* c, const double* d_a, uint32_t l, double* d_evals, double* d_evecs,
                       const int* d_skip_flag = nullptr);
// CholeskyQR transform T = R^-1 with G = R^T R; *d_ok_flag = 1 on success, 0 when a pivot <= rel_eps * max diag
int launch_chol_orth(gpca_ctx* c, const double* d_g, uint32_t l, double* d_t, double rel_eps, int* d_ok_flag);
// T = V * diag(lambda^-1/2) with rank truncation (lambda <= eps*lambda_max -> column zeroed)
int launch_make_orth_transform(gpca_ctx* c, const double* d_evals, const double* d_evecs, uint32_t l, double* d_t,
                               double rel_eps, const int* d_skip_flag = nullptr);
int launch_f32_to_f64(gpca_ctx* c, const float* in, double* out, uint64_t n);

// ---- batched dense helpers: the same operations for many small problems (one per LD block) in one launch ----------
// Problem b is the [rows x l] fp32 matrix at base + off with the shared row stride ld (l <= 32).
struct DenseProb {
  uint64_t off;
  uint32_t rows;
  uint32_t l;
};
struct DenseBatchWs {   // device scratch for n problems
  double* G;       // [n][1024]  l x l Gram, compact
  double* T;       // [n][1024]  transform
  double* evecs;   // [n][1024]
  double* evals;   // [n][32]
  int* flags;      // [n]        1 = Cholesky transform valid
};
int dense_batch_ws(gpca_ctx* c, uint32_t n_probs, DenseBatchWs& ws);
// out rows keyed by (seed, d_streams[b], r, col); columns >= l zeroed up to ld
int launch_gaussian_batch(gpca_ctx* c, float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                          uint64_t max_rows, uint64_t seed, const uint32_t* d_streams);
// Gram partials of every problem (into c->ws_gram); *nparts = partials per problem
int launch_gram_batch(gpca_ctx* c, const float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                      uint64_t max_rows, int* nparts);
// sums the partials -> ws.G, Cholesky transform -> ws.T, ws.flags
int launch_chol_orth_batch(gpca_ctx* c, int nparts, const DenseProb* d_probs, uint32_t n_probs, double rel_eps,
                           const DenseBatchWs& ws);
// eigen-decomposition of ws.G -> ws.evals / ws.evecs (problems whose flag is set are skipped when use_flags)
int launch_jacobi_eigh_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, const DenseBatchWs& ws,
                             bool use_flags);
int launch_make_orth_transform_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, double rel_eps,
                                     const DenseBatchWs& ws);
// ws.T_b [l x l2_b] = ws.evecs_b[:, :l2_b]
int launch_rotation_batch(gpca_ctx* c, const DenseProb* d_probs, uint32_t n_probs, const uint32_t* d_l2,
                          const DenseBatchWs& ws);
// out_b = Y_b * T_b  (T_b [l x l2_b]; d_l2 == nullptr: l2_b = l).  d_out_offs == nullptr: out_b at d_out + off (in
// place allowed when d_out == d_base and ldo == ld).
int launch_apply_right_batch(gpca_ctx* c, const float* d_base, uint32_t ld, const DenseProb* d_probs,
                             uint32_t n_probs, uint64_t max_rows, const double* d_t, const uint32_t* d_l2,
                             uint32_t max_l2, float* d_out, const uint64_t* d_out_offs, uint32_t ldo);
// two rounds of (Gram, CholeskyQR with guarded eigen fallback, apply), every problem in place
int orthonormalize_batch(gpca_ctx* c, float* d_base, uint32_t ld, const DenseProb* d_probs, uint32_t n_probs,
                         uint64_t max_rows, uint32_t max_l, const DenseBatchWs& ws);
int launch_scale_cols_to_f64(gpca_ctx* c, const float* in, uint64_t n, uint32_t k, uint32_t ld, double* out);

// ---- sketch (kernels_sketch.cu) -----------------------------------------------------------
// bound on |z| of the normals of philox_normal4 (philox.cuh): u1 >= 2^-25 -> radius <= sqrt(50 ln 2) = 5.887
#define GPCA_NORMAL_ABS_MAX 5.9f

struct SketchProblem {
  PackedMat G;            // [rows x K]
  const float* Bin;       // [K x ld] dense operand
  uint32_t l, ld;         // logical columns (<=64), row stride
  const float* f;         // [K] per-k operand scale (nullptr = 1)
  const float* e;         // [K] per-k weight for the rank-one term (nullptr = 1)
  const float* a;         // [rows] per-row output scale (nullptr = 1)
  const float* b;         // [rows] per-row weight of the rank-one term (nullptr = 1)
  float* out;             // [rows x ldo]
  uint32_t ldo;
  bool out_pad = false;   // columns [l, ldo) of `out` are padding the pass may overwrite with zeros (wide row stores)
  // Operand statistics as by-products (integer engine; ignored elsewhere).  A sample-side pass needs, for its operand
  // W, the column sums e^T W and max |f o W|; when W was just produced by a snp-side pass (f = that pass's a, e = its
  // b) or by the Gaussian generator, the producer computes them on the way out and one sweep over W is saved.
  // Generated operand: Bin[k, n] = the standard normal that philox_normal4(gen_seed, gen_stream, gen_row0 + k, n/4)
  // gives for column n (the test matrix Omega).  The integer engine quantises it straight from the generator -- the
  // K x ld fp32 matrix is never written or read -- with the a-priori scale gen_amax >= max |f o Bin|; every other
  // path fills `Bin` (which then must be writable scratch of K x ld floats) first.
  bool gen = false;
  uint64_t gen_seed = 0, gen_row0 = 0;
  uint32_t gen_stream = 0;
  float gen_amax = 0.f;
  bool emit_stats = false;   // also leave the statistics of `out` (weights a, b) in the context
  bool use_stats = false;    // Bin's statistics are in the context already
};
// out[r,:] = a_r * sum_k code(r,k) f_k Bin[k,:]  -  b_r * sum_k e_k Bin[k,:]  (+ missing correction)
int launch_sketch(gpca_ctx* c, const SketchProblem& p);

// sign convention helpers (kernels_dense.cu): largest-|.| entry of every column positive
int launch_sign_flags(gpca_ctx* c, const float* d_x, uint64_t n, uint32_t k, uint32_t ld, int* d_flags);
// out[i][j] = flags[j] ? -x[i][j] : x[i][j], compact [n x k], as f32 and/or f64 (either may be null)
int launch_apply_flags(gpca_ctx* c, const float* d_x, uint64_t n, uint32_t k, uint32_t ld, const int* d_flags,
                       float* d_out_f32, double* d_out_f64);
