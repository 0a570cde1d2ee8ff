// drivers.cu -- the rfit PCA driver built from the sketch passes, plus helpers shared with eigensnp.cu.
//   gpca_rfit replaces pca_runner::run_genomic_pca = PCA::rfit + PCA::transform (src/main.rs:598-679)
// The arithmetic lives in the external efficient_pca crate (parity unpinned); the stage order implemented
// here is the one restated in oracle/pca.py::rfit, which the tests check it against.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <random>

#include "driver_util.cuh"
#include "parallel_for.h"

static int fail(gpca_ctx* c, int code, const std::string& msg) {
  c->set_error(msg);
  return code;
}

constexpr uint32_t STREAM_RFIT_OMEGA = 1;

int get_small(gpca_ctx* c, Small& s) {
  GPCA_CUDA_TRY(c, c->ws_small.alloc(4 * 64 * 64 + 8));
  s.G = c->ws_small.p;
  s.evals = s.G + 64 * 64;
  s.evecs = s.evals + 64 * 64;
  s.T = s.evecs + 64 * 64;
  s.flag = reinterpret_cast<int*>(s.T + 64 * 64);
  return GPCA_OK;
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, uint64_t n) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x)
    out[t] = (double)in[t];
}
int launch_f32_to_f64(gpca_ctx* c, const float* in, double* out, uint64_t n) {
  if (n == 0) return GPCA_OK;
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
  f32_to_f64_kernel<<<grid, 256, 0, c->stream>>>(in, out, n);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  return GPCA_OK;
}

int download_results(gpca_ctx* c, const float* d_src, uint64_t count, float* dst_f32, double* dst_f64) {
  constexpr uint64_t CH_MAX = 1u << 21;      // floats per chunk (8 MB landing buffers)
  if (count == 0 || (!dst_f32 && !dst_f64)) return GPCA_OK;
  {
    // a caller-owned buffer that is pinned (cudaHostRegister / gpca_host_alloc) takes the results straight off the bus:
    // f32 as they are, f64 widened on the device first
    void* dst = dst_f32 ? (void*)dst_f32 : (void*)dst_f64;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();      // (an unregistered pointer is not an error here)
    if (pinned && !getenv("GPCA_DEBUG_DOWNLOAD_CHUNK")) {
      if (dst_f32) {
        GPCA_CUDA_TRY(c, cudaMemcpyAsync(dst_f32, d_src, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      } else {
        GPCA_CUDA_TRY(c, c->ws_f64.alloc(count));
        GPCA_TRY(launch_f32_to_f64(c, d_src, c->ws_f64.p, count));
        GPCA_CUDA_TRY(c, cudaMemcpyAsync(dst_f64, c->ws_f64.p, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      }
      GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      return GPCA_OK;
    }
  }
  uint64_t CH = CH_MAX;
  if (const char* e = getenv("GPCA_DEBUG_DOWNLOAD_CHUNK")) {   // tests: force many chunks on small results
    const uint64_t v = strtoull(e, nullptr, 10);
    if (v >= 1 && v <= CH_MAX) CH = v;
  }
  for (int i = 0; i < 2; ++i) {
    if (!c->h_dl[i]) GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_dl[i], CH_MAX * sizeof(float)));
    if (!c->ev_dl[i]) GPCA_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_dl[i], cudaEventDisableTiming));
  }
  auto land = [&](uint64_t q) {          // chunk q has arrived in its buffer: hand it to the caller's memory
    const uint64_t off = q * CH, cnt = std::min<uint64_t>(CH, count - off);
    const float* src = c->h_dl[q & 1];
    parallel_for(cnt, [&](uint64_t lo, uint64_t hi) {
      if (dst_f32) std::memcpy(dst_f32 + off + lo, src + lo, (hi - lo) * sizeof(float));
      if (dst_f64)
        for (uint64_t i = lo; i < hi; ++i) dst_f64[off + i] = (double)src[i];
    }, 1u << 17);
  };
  const uint64_t nch = (count + CH - 1) / CH;
  for (uint64_t q = 0; q < nch; ++q) {
    const uint64_t off = q * CH, cnt = std::min<uint64_t>(CH, count - off);
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->h_dl[q & 1], d_src + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GPCA_CUDA_TRY(c, cudaEventRecord(c->ev_dl[q & 1], c->stream));
    if (q > 0) {
      GPCA_CUDA_TRY(c, cudaEventSynchronize(c->ev_dl[(q - 1) & 1]));
      land(q - 1);
    }
  }
  GPCA_CUDA_TRY(c, cudaEventSynchronize(c->ev_dl[(nch - 1) & 1]));
  land(nch - 1);
  return GPCA_OK;
}

int orthonormalize(gpca_ctx* c, float* y, uint64_t n, uint32_t l, uint32_t ld, bool sharded, const Small& s) {
  for (int rep = 0; rep < 2; ++rep) {
    GPCA_TRY(launch_gram(c, y, n, l, ld, s.G));
    if (sharded) GPCA_TRY(driver_allreduce(c, s.G, (uint64_t)l * l, 1));
    // CholeskyQR first (tens of microseconds); the eigen-based transform runs only if a pivot was unsafe
    const double eps = rep == 0 ? 1e-11 : 1e-13;
    GPCA_TRY(launch_chol_orth(c, s.G, l, s.T, eps, s.flag));
    GPCA_TRY(launch_jacobi_eigh(c, s.G, l, s.evals, s.evecs, s.flag));
    GPCA_TRY(launch_make_orth_transform(c, s.evals, s.evecs, l, s.T, eps, s.flag));
    GPCA_TRY(launch_apply_right(c, y, n, l, ld, s.T, l, y, ld));
  }
  return GPCA_OK;
}

__global__ void rotation_transform_kernel(const double* evals, const double* evecs, uint32_t l, uint32_t k, double* t,
                                          int inv_sqrt) {
  for (int i = threadIdx.x; i < (int)(l * k); i += blockDim.x) {
    const int r = i / k, j = i % k;
    const double lam = evals[j];
    const double v = evecs[r * l + j];
    t[i] = inv_sqrt ? (lam > 0.0 ? v / sqrt(lam) : 0.0) : v;
  }
}

int launch_rotation_transform(gpca_ctx* c, const double* evals, const double* evecs, uint32_t l, uint32_t k, double* t,
                              bool inv_sqrt) {
  rotation_transform_kernel<<<1, 256, 0, c->stream>>>(evals, evecs, l, k, t, inv_sqrt ? 1 : 0);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  return GPCA_OK;
}

void fix_signs_host(std::vector<float>& scores, uint64_t n, uint32_t k, std::vector<int>& flip) {
  flip.assign(k, 0);
  for (uint32_t j = 0; j < k; ++j) {
    float best = -1.f;
    uint64_t bi = 0;
    for (uint64_t i = 0; i < n; ++i) {
      const float a = std::fabs(scores[i * k + j]);
      if (a > best) {
        best = a;
        bi = i;
      }
    }
    if (n && scores[bi * k + j] < 0) {
      flip[j] = 1;
      for (uint64_t i = 0; i < n; ++i) scores[i * k + j] = -scores[i * k + j];
    }
  }
}

extern "C" int gpca_rfit(gpca_ctx* c, uint32_t k, uint32_t oversample, uint32_t power_iters, uint64_t seed,
                         int has_seed, double* scores, double* eigenvalues, float* loadings, uint32_t* k_out) {
  if (!c) return GPCA_ERR_INVALID;
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  GPCA_HOST_POOL(c);
  if (c->D == 0) return fail(c, GPCA_ERR_INVALID, "gpca_set_pca_snps has not been called");
  if (k == 0) return fail(c, GPCA_ERR_INVALID, "Number of components (-k) must be > 0.");  // main.rs:607
  const uint64_t N = c->N, D = c->D;
  const uint64_t Dtot = c->shard_total ? c->shard_total : D;
  if (N < 2) return fail(c, GPCA_ERR_INVALID, "PCA requires at least 2 samples");            // main.rs:614
  const uint64_t maxk = std::min<uint64_t>(N, Dtot);                                        // main.rs:621
  if (k > maxk) k = (uint32_t)maxk;                                                         // main.rs:622-628
  uint64_t l64 = std::min<uint64_t>((uint64_t)k + oversample, maxk);
  if (l64 > 64) return fail(c, GPCA_ERR_INVALID, "k + oversample must be <= 64 in this build");
  const uint32_t l = (uint32_t)l64;
  Small s;
  GPCA_TRY(get_small(c, s));
  if (!has_seed) {
    std::random_device rd;
    seed = ((uint64_t)rd() << 32) ^ rd() ^ (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count();
    if (c->sharded()) {
      // every shard must sketch with the same test matrix: shard 0's entropy seed goes to all of them.  The host hook
      // has no broadcast, so there an explicit seed is required.
      if (!c->nccl_comm)
        return fail(c, GPCA_ERR_INVALID, "sharded gpca_rfit through the allreduce hook needs an explicit seed (has_seed = 1)");
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(s.G, &seed, 8, cudaMemcpyHostToDevice, c->stream));
      GPCA_TRY(driver_broadcast0(c, s.G, 8));
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(&seed, s.G, 8, cudaMemcpyDeviceToHost, c->stream));
      GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
  }
  DevBuf<float>&Y = c->drv_a, &Z = c->drv_b, &R = c->drv_c, &Sc = c->drv_d;
  // rows of the D x l matrices are padded to a multiple of 8 floats: the snp-side sketch writes a row per thread, and
  // 32-byte aligned rows let it use 32-byte stores (one full sector per store instead of four partial writes)
  const uint32_t ldz = (l + 7u) & ~7u;
  GPCA_CUDA_TRY(c, Y.alloc(N * l));
  GPCA_CUDA_TRY(c, Z.alloc(D * ldz));

  // Y = S^T Omega
  // (every D x l operand of a sample-side pass arrives with its column statistics: from the generator here, from the
  //  epilogue of the snp-side pass below -- one sweep over a D x l matrix saved per pass)
  // (Omega is quantised straight from the generator when the integer engine runs: the D x l fp32 matrix is neither
  //  written nor read back -- a 1.3 GB write and read at config 3)
  GPCA_TRY(sketch_sample_side_gaussian(c, Z.p, Y.p, l, ldz, l, seed, STREAM_RFIT_OMEGA));
  // One side is re-orthonormalised per iteration, the other is left as it comes (range(S^T Z) does not depend on a
  // column transform of Z, nor range(S Y) on one of Y): the side with fewer rows on this device -- the samples at the
  // 1000G shape, the shard's SNPs at the 500,000-sample shapes (the l x l Gram of a sharded D side is summed over the
  // shards through the allreduce hook).
  // Sharded, the choice must be the same on every shard (it decides whether the l x l Gram allreduce is issued): it is
  // made from the MEAN shard size, which every shard learns from one two-number allreduce (cached per shard size).
  bool orth_snp_side = D < N;
  if (c->sharded()) {
    if (c->shard_vote_snp_side < 0 || c->shard_vote_D != D) {
      double h[2] = {(double)D, 1.0};
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(s.G, h, 16, cudaMemcpyHostToDevice, c->stream));
      GPCA_TRY(driver_allreduce(c, s.G, 2, 1));
      GPCA_CUDA_TRY(c, cudaMemcpyAsync(h, s.G, 16, cudaMemcpyDeviceToHost, c->stream));
      GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      c->shard_vote_snp_side = (h[1] > 0.0 && h[0] / h[1] < (double)N) ? 1 : 0;
      c->shard_vote_D = D;
    }
    orth_snp_side = c->shard_vote_snp_side == 1;
  }
  if (getenv("GPCA_DEBUG_ORTH_SAMPLE_SIDE")) orth_snp_side = false;
  for (uint32_t it = 0; it < power_iters; ++it) {
    if (!orth_snp_side) {
      GPCA_TRY(orthonormalize(c, Y.p, N, l, l, false, s));
      GPCA_TRY(sketch_snp_side(c, Y.p, Z.p, l, l, ldz, true, true));   // Z = S Q
      GPCA_TRY(sketch_sample_side(c, Z.p, Y.p, l, ldz, l, true));      // Y = S^T Z
    } else {
      GPCA_TRY(sketch_snp_side(c, Y.p, Z.p, l, l, ldz, false, true));  // Z = S Y
      GPCA_TRY(orthonormalize(c, Z.p, D, l, ldz, c->sharded(), s));
      c->stats_for = nullptr;                                          // (any statistics of Z are stale now)
      GPCA_TRY(sketch_sample_side(c, Z.p, Y.p, l, ldz, l, false));     // Y = S^T Q_z
    }
  }
  GPCA_TRY(orthonormalize(c, Y.p, N, l, l, false, s));
  GPCA_TRY(sketch_snp_side(c, Y.p, Z.p, l, l, ldz, true, true));  // B = S Q   [D x l]
  // B^T B = Q^T (S^T B): the l x l matrix whose eigen-decomposition gives the singular values and the right factor of
  // B comes from the sample-side sketch of B -- which is also all that the scores need (scores = S^T B V_b / s).  The
  // D-row Gram of B and, when the rotation is not asked for, every D x l by l x k product disappear; on several GPUs
  // S^T B is already summed over the shards, so no further exchange is needed.
  DevBuf<float>& Y2 = c->drv_e;
  GPCA_CUDA_TRY(c, Y2.alloc(N * l));
  GPCA_TRY(sketch_sample_side(c, Z.p, Y2.p, l, ldz, l, true));    // S^T B   [N x l]
  GPCA_TRY(launch_cross_gram(c, Y.p, Y2.p, N, l, l, s.G));
  GPCA_TRY(launch_jacobi_eigh(c, s.G, l, s.evals, s.evecs));
  GPCA_TRY(launch_rotation_transform(c, s.evals, s.evecs, l, k, s.T, true));
  GPCA_CUDA_TRY(c, Sc.alloc(N * k));
  GPCA_TRY(launch_apply_right(c, Y2.p, N, l, l, s.T, k, Sc.p, k));   // transform(): scores = S^T rotation
  if (loadings) {
    GPCA_CUDA_TRY(c, R.alloc(D * k));
    GPCA_TRY(launch_apply_right(c, Z.p, D, l, ldz, s.T, k, R.p, k));   // rotation = B V_b / s  [D x k]
  }

  // sign convention, f32 -> f64 conversion and the flips of the loadings all happen on the device;
  // the host only receives the final buffers
  std::vector<double> h_ev(l);
  DevBuf<int> d_flags;
  GPCA_CUDA_TRY(c, d_flags.alloc(64));
  GPCA_TRY(launch_sign_flags(c, Sc.p, N, k, k, d_flags.p));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(h_ev.data(), s.evals, l * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (scores) {
    // fp32 on the bus (the scores are computed in fp32; widening to the f64 the reference's API returns is exact and
    // happens on host threads behind the transfer)
    GPCA_TRY(launch_apply_flags(c, Sc.p, N, k, k, d_flags.p, Sc.p, nullptr));
    GPCA_TRY(download_results(c, Sc.p, N * k, nullptr, scores));
  }
  if (loadings) {
    GPCA_TRY(launch_apply_flags(c, R.p, D, k, k, d_flags.p, R.p, nullptr));
    GPCA_TRY(download_results(c, R.p, D * k, loadings, nullptr));
  }
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (eigenvalues)
    for (uint32_t j = 0; j < k; ++j) eigenvalues[j] = h_ev[j] / (double)(N - 1);
  if (k_out) *k_out = k;
  return GPCA_OK;
}

extern "C" void gpca_eigensnp_default_cfg(gpca_eigensnp_cfg* cfg) {
  if (!cfg) return;
  cfg->target_num_global_pcs = 10;      // main.rs:554
  cfg->components_per_ld_block = 7;     // main.rs:557
  cfg->subset_factor = 0.075;           // main.rs:560
  cfg->min_subset_size = 10000;         // main.rs:563
  cfg->max_subset_size = 40000;         // main.rs:566
  cfg->global_oversampling = 10;        // main.rs:569
  cfg->global_power_iters = 2;          // main.rs:572
  cfg->local_oversampling = 10;         // main.rs:575
  cfg->local_power_iters = 2;           // main.rs:578
  cfg->random_seed = 2025;              // main.rs:581
  cfg->snp_processing_strip_size = 2000;  // main.rs:584
  cfg->refine_pass_count = 1;           // main.rs:587
  cfg->collect_diagnostics = 0;
}
