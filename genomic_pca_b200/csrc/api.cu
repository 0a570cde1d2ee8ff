// api.cu -- the C ABI (include/gpca.h): context, ingest, statistics, sketch entry points.
// The PCA drivers (rfit, EigenSNP) live in drivers.cu.
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cmath>
#include <cstring>
#include <atomic>
#include <new>

#include "host_qc.h"
#include "kernels.cuh"
#include "sketch_tc.cuh"
#include "parallel_for.h"

#define CHECK_CTX(c) \
  if (!(c)) return GPCA_ERR_INVALID;

void gpca_destroy_cublas(void* h);   // eigensnp.cu

static int fail(gpca_ctx* c, int code, const std::string& msg) {
  c->set_error(msg);
  return code;
}

extern "C" const char* gpca_version(void) { return "genomic_pca_b200 0.1 (sm_100a)"; }

extern "C" int gpca_init(gpca_ctx** out, int device) {
  if (!out) return GPCA_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return GPCA_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GPCA_ERR_NO_DEVICE;
  if (prop.major != 10) return GPCA_ERR_NO_DEVICE;  // sm_100a binary only: no fallback of any kind
  if (cudaSetDevice(device) != cudaSuccess) return GPCA_ERR_CUDA;
  gpca_ctx* c = new (std::nothrow) gpca_ctx();
  if (!c) return GPCA_ERR_OOM;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return GPCA_ERR_CUDA;
  }
  const char* eng = getenv("GPCA_SKETCH_ENGINE");
  if (eng) c->engine = atoi(eng);
  const char* bb = getenv("GPCA_BATCH_BLOCKS");
  if (bb) c->batch_blocks = atoi(bb) != 0;
  *out = c;
  return GPCA_OK;
}

extern "C" void gpca_destroy(gpca_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& pr : c->pending_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (auto& pr : c->pending_kernel_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  if (c->h_up) cudaFreeHost(c->h_up);
  for (int i = 0; i < 2; ++i) {
    if (c->h_dl[i]) cudaFreeHost(c->h_dl[i]);
    if (c->ev_dl[i]) cudaEventDestroy(c->ev_dl[i]);
  }
  for (int i = 0; i < 2; ++i)
    if (c->h_rd[i]) cudaFreeHost(c->h_rd[i]);
  gpca_destroy_cublas(c->cublas);
  delete c;
}

extern "C" const char* gpca_last_error(const gpca_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" uint64_t gpca_launch_count(const gpca_ctx* c) { return c ? c->launches : 0; }
extern "C" void gpca_reset_launch_count(gpca_ctx* c) {
  if (c) c->launches = 0;
}
extern "C" int gpca_set_sketch_engine(gpca_ctx* c, int engine) {
  CHECK_CTX(c);
  if (engine < 0 || engine > 2) return fail(c, GPCA_ERR_INVALID, "engine must be 0 (SIMT), 1 (tcgen05 f16) or 2 (tcgen05 i8)");
  c->engine = engine;
  return GPCA_OK;
}
extern "C" int gpca_set_batch_blocks(gpca_ctx* c, int on) {
  CHECK_CTX(c);
  c->batch_blocks = on ? 1 : 0;
  return GPCA_OK;
}
extern "C" int gpca_set_allreduce(gpca_ctx* c, gpca_allreduce_fn fn, void* user) {
  CHECK_CTX(c);
  c->allreduce = fn;
  c->allreduce_user = user;
  return GPCA_OK;
}
extern "C" int gpca_set_shard(gpca_ctx* c, uint64_t off, uint64_t total) {
  CHECK_CTX(c);
  c->shard_offset = off;
  c->shard_total = total;
  return GPCA_OK;
}
extern "C" uint64_t gpca_num_samples(const gpca_ctx* c) { return c ? c->N : 0; }
extern "C" uint64_t gpca_num_snps(const gpca_ctx* c) { return c ? c->M : 0; }
extern "C" uint64_t gpca_num_pca_snps(const gpca_ctx* c) { return c ? c->D : 0; }
extern "C" void* gpca_get_stream(gpca_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int gpca_synchronize(gpca_ctx* c) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return GPCA_OK;
}

static void reset_loaded(gpca_ctx* c) {
  c->have_counts = false;   // (the host vectors keep their storage: re-growing them would zero-fill hundreds of MB)
  c->D = 0;
  c->Gs = PackedMat();
  c->Gt = PackedMat();   // (the stores are kept: DevBuf::alloc reuses them when the next data set fits)
  c->any_missing = false;
  c->es_store.release();
  c->et_store.release();
  c->ets_store.release();
  c->ess_store.release();
  c->es_cn.release();
  c->es_pool.release();
}

// ---- ingest ------------------------------------------------------------------------------
static int alloc_raw(gpca_ctx* c, uint64_t N, uint64_t M) {
  reset_loaded(c);
  c->N = N;
  c->M = M;
  c->raw_pitch = round_up((N + 3) / 4, 16);
  GPCA_CUDA_TRY(c, c->raw.alloc(std::max<size_t>(c->raw_pitch * M, 16)));
  return GPCA_OK;
}

extern "C" int gpca_load_bed(gpca_ctx* c, const uint8_t* host_payload, uint64_t n_in, uint64_t n_snps,
                             const int64_t* keep, uint64_t n_keep) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!host_payload && n_snps) return fail(c, GPCA_ERR_INVALID, "null payload");
  if (n_in == 0) return fail(c, GPCA_ERR_INVALID, "no samples");
  const uint64_t N = keep ? n_keep : n_in;
  if (N == 0) return fail(c, GPCA_ERR_INVALID, "No samples passed QC.");  // prepare.rs:1010
  if (keep)
    for (uint64_t i = 0; i < n_keep; ++i)
      if (keep[i] < 0 || (uint64_t)keep[i] >= n_in || (i && keep[i] <= keep[i - 1]))
        return fail(c, GPCA_ERR_INVALID, "keep_samples must be increasing indices into the FAM order");
  c->vcf_mode = false;
  GPCA_TRY(alloc_raw(c, N, n_snps));
  DevBuf<int64_t> d_keep;
  if (keep) {
    GPCA_CUDA_TRY(c, d_keep.alloc(n_keep));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_keep.p, keep, n_keep * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  }
  const size_t in_pitch = (n_in + 3) / 4;
  // stream the payload through two staging buffers (rows per chunk sized to ~256 MiB)
  uint64_t rows_per_chunk = std::max<uint64_t>(1, (256ull << 20) / std::max<size_t>(in_pitch, 1));
  rows_per_chunk = std::min<uint64_t>(rows_per_chunk, std::max<uint64_t>(n_snps, 1));
  DevBuf<uint8_t> stage[2];
  cudaEvent_t done[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    GPCA_CUDA_TRY(c, stage[i].alloc(rows_per_chunk * in_pitch + 16));
    GPCA_CUDA_TRY(c, cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  int rc = GPCA_OK;
  int buf = 0;
  for (uint64_t r0 = 0; r0 < n_snps && rc == GPCA_OK; r0 += rows_per_chunk, buf ^= 1) {
    const uint64_t nr = std::min<uint64_t>(rows_per_chunk, n_snps - r0);
    cudaError_t e = cudaEventSynchronize(done[buf]);  // previous use of this staging buffer finished
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(stage[buf].p, host_payload + r0 * in_pitch, nr * in_pitch, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      rc = fail(c, GPCA_ERR_CUDA, std::string("load_bed H2D: ") + cudaGetErrorString(e));
      break;
    }
    rc = launch_repitch_gather(c, stage[buf].p, in_pitch, n_in, keep ? d_keep.p : nullptr, N, nr,
                               c->raw.p + r0 * c->raw_pitch, c->raw_pitch);
    cudaEventRecord(done[buf], c->stream);
  }
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]);
  return rc;
}

extern "C" int gpca_load_bed_device(gpca_ctx* c, const uint8_t* dev_payload, uint64_t n_samples, uint64_t n_snps) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (n_samples == 0) return fail(c, GPCA_ERR_INVALID, "no samples");
  c->vcf_mode = false;
  GPCA_TRY(alloc_raw(c, n_samples, n_snps));
  GPCA_TRY(launch_repitch_gather(c, dev_payload, (n_samples + 3) / 4, n_samples, nullptr, n_samples, n_snps, c->raw.p,
                                 c->raw_pitch));
  return GPCA_OK;
}

extern "C" int gpca_load_u8_variant_major(gpca_ctx* c, const uint8_t* host_dosage, uint64_t n_samples,
                                          uint64_t n_variants) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (n_samples == 0) return fail(c, GPCA_ERR_INVALID, "No samples available to build matrix.");  // vcf.rs:325
  if (n_variants == 0) return fail(c, GPCA_ERR_INVALID, "No variants available to build matrix.");  // vcf.rs:322
  GPCA_TRY(alloc_raw(c, n_samples, n_variants));
  c->vcf_mode = true;
  uint64_t rows_per_chunk = std::max<uint64_t>(1, (256ull << 20) / n_samples);
  rows_per_chunk = std::min<uint64_t>(rows_per_chunk, n_variants);
  DevBuf<uint8_t> stage;
  GPCA_CUDA_TRY(c, stage.alloc(rows_per_chunk * n_samples));
  for (uint64_t r0 = 0; r0 < n_variants; r0 += rows_per_chunk) {
    const uint64_t nr = std::min<uint64_t>(rows_per_chunk, n_variants - r0);
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(stage.p, host_dosage + r0 * n_samples, nr * n_samples, cudaMemcpyHostToDevice,
                                     c->stream));
    GPCA_TRY(launch_u8_to_plink(c, stage.p, n_samples, nr, c->raw.p + r0 * c->raw_pitch, c->raw_pitch));
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  return GPCA_OK;
}

// ---- statistics ----------------------------------------------------------------------------
static int ensure_counts(gpca_ctx* c) {
  if (c->have_counts) return GPCA_OK;
  if (!c->raw.p) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded (or it was released)");
  const uint64_t M = c->M;
  GPCA_CUDA_TRY(c, c->d_cnt.alloc(std::max<uint64_t>(M, 1)));
  GPCA_TRY(launch_bed_counts(c, c->raw.p, c->raw_pitch, M, c->d_cnt.p));
  if (c->h_cnt_cap < M) {   // pinned landing buffer for the 16 B/SNP count records, kept across calls
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    c->h_cnt = nullptr;
    c->h_cnt_cap = 0;
    GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_cnt, std::max<uint64_t>(M, 1) * sizeof(uint4)));
    c->h_cnt_cap = M;
  }
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->h_cnt, c->d_cnt.p, M * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const uint32_t pad = (uint32_t)(c->raw_pitch * 4 - c->N);  // pad fields were written as 01
  c->h_counts.resize(M * 4);
  uint32_t* hc = c->h_counts.data();
  const uint32_t n32 = (uint32_t)c->N;
  const uint4* hp = c->h_cnt;
  parallel_for(M, [=](uint64_t lo, uint64_t hi) {
    for (uint64_t j = lo; j < hi; ++j) {
      const uint32_t miss = hp[j].x - pad, het = hp[j].y, d0 = hp[j].z;
      const uint32_t nv = n32 - miss;
      hc[4 * j + 0] = nv;
      hc[4 * j + 1] = d0;             // code 11 -> dosage 0
      hc[4 * j + 2] = het;            // code 10 -> dosage 1
      hc[4 * j + 3] = nv - d0 - het;  // code 00 -> dosage 2
    }
  });
  c->have_counts = true;
  return GPCA_OK;
}

extern "C" int gpca_snp_counts(gpca_ctx* c, uint32_t* n_valid, uint32_t* n0, uint32_t* n1, uint32_t* n2) {
  CHECK_CTX(c);
  GPCA_TRY(ensure_counts(c));
  for (uint64_t j = 0; j < c->M; ++j) {
    if (n_valid) n_valid[j] = c->h_counts[4 * j + 0];
    if (n0) n0[j] = c->h_counts[4 * j + 1];
    if (n1) n1[j] = c->h_counts[4 * j + 2];
    if (n2) n2[j] = c->h_counts[4 * j + 3];
  }
  return GPCA_OK;
}

extern "C" int gpca_snp_qc(gpca_ctx* c, const gpca_qc_cfg* cfg, uint8_t* keep, float* mean, float* sd,
                           uint8_t* fail_code) {
  CHECK_CTX(c);
  if (!cfg || !keep) return fail(c, GPCA_ERR_INVALID, "cfg/keep null");
  GPCA_TRY(ensure_counts(c));
  host_snp_qc(c->N, c->M, c->h_counts.data(), *cfg, keep, mean, sd, fail_code);
  return GPCA_OK;
}

extern "C" int gpca_vcf_maf_filter(gpca_ctx* c, double maf_threshold, uint8_t* keep, float* mean, float* sd) {
  CHECK_CTX(c);
  if (!keep) return fail(c, GPCA_ERR_INVALID, "keep null");
  GPCA_TRY(ensure_counts(c));
  host_vcf_maf(c->N, c->M, c->h_counts.data(), maf_threshold, keep, mean, sd);
  return GPCA_OK;
}

// ---- PCA SNP set -----------------------------------------------------------------------------
static int build_pca_set(gpca_ctx* c, uint64_t D) {
  // c->pca_idx, c->h_mean, c->h_sd are filled; derive the device-side vectors and the resident copies
  const uint64_t* idx = c->pca_idx.data();
  const float* mean = c->h_mean.data();
  const float* sd = c->h_sd.data();
  c->h_inv.resize(D);
  c->h_muinv.resize(D);
  float* inv = c->h_inv.data();
  float* muinv = c->h_muinv.data();
  const uint32_t* hc = c->h_counts.data();
  const uint64_t N = c->N;
  std::atomic<uint64_t> nmiss_total{0};
  std::atomic<uint32_t> inv_max_bits{0};     // (bit patterns of non-negative floats order like the floats)
  parallel_for(D, [&, idx, mean, sd, inv, muinv, hc, N](uint64_t lo, uint64_t hi) {
    uint64_t nm = 0;
    float imax = 0.f;
    for (uint64_t i = lo; i < hi; ++i) {
      // same f32 expressions as prepare.rs:1948-1949 (recip, mean*recip); sd < 1e-9 -> the row standardises to 0
      if (std::fabs(sd[i]) < 1e-9f) {
        inv[i] = 0.f;
        muinv[i] = 0.f;
      } else {
        const float r = 1.0f / sd[i];
        inv[i] = r;
        muinv[i] = mean[i] * r;
      }
      nm += N - hc[4 * idx[i]];
      imax = std::max(imax, std::fabs(inv[i]));
    }
    nmiss_total.fetch_add(nm);
    uint32_t bits, cur = inv_max_bits.load();
    std::memcpy(&bits, &imax, 4);
    while (bits > cur && !inv_max_bits.compare_exchange_weak(cur, bits)) {
    }
  });
  {
    const uint32_t bits = inv_max_bits.load();
    std::memcpy(&c->inv_sd_max, &bits, 4);
  }
  c->any_missing = nmiss_total.load() > 0;
  c->D = D;
  GPCA_CUDA_TRY(c, c->d_mean.alloc(D));
  GPCA_CUDA_TRY(c, c->d_sd.alloc(D));
  GPCA_CUDA_TRY(c, c->d_inv_sd.alloc(D));
  GPCA_CUDA_TRY(c, c->d_mu_inv_sd.alloc(D));
  GPCA_CUDA_TRY(c, c->d_idx.alloc(D));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->d_idx.p, idx, D * 8, cudaMemcpyHostToDevice, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->d_mean.p, mean, D * 4, cudaMemcpyHostToDevice, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->d_sd.p, sd, D * 4, cudaMemcpyHostToDevice, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->d_inv_sd.p, inv, D * 4, cudaMemcpyHostToDevice, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->d_mu_inv_sd.p, muinv, D * 4, cudaMemcpyHostToDevice, c->stream));
  c->Gs.rows = D;
  c->Gs.cols = c->N;
  c->Gs.pitch = round_up((c->N + 3) / 4, 128);
  c->Gt.rows = c->N;
  c->Gt.cols = D;
  c->Gt.pitch = round_up((D + 3) / 4, 128);
  GPCA_CUDA_TRY(c, c->gs_store.alloc(c->Gs.pitch * D));
  c->Gs.p = c->gs_store.p;
  GPCA_TRY(launch_build_gs(c, c->raw.p, c->raw_pitch, c->d_idx.p, c->Gs));
  // the PLINK-coded staging copy is only needed again for a different SNP selection; under memory pressure
  // (three copies would not fit comfortably) it is released before the transposed copy is allocated
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (c->gt_store.n < c->Gt.pitch * c->N && free_b < c->Gt.pitch * c->N + (8ull << 30)) {
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->raw.release();
  }
  GPCA_CUDA_TRY(c, c->gt_store.alloc(c->Gt.pitch * c->N));
  c->Gt.p = c->gt_store.p;
  GPCA_TRY(launch_transpose(c, c->Gs, c->Gt));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return GPCA_OK;
}

extern "C" int gpca_set_pca_snps(gpca_ctx* c, const uint64_t* snp_idx, uint64_t D, const float* mean,
                                 const float* sd) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->raw.p) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded (or it was released)");
  if (D == 0) return fail(c, GPCA_ERR_INVALID, "No SNPs passed all QC filters.");  // prepare.rs:1020
  if (!snp_idx || !mean || !sd) return fail(c, GPCA_ERR_INVALID, "null argument");
  std::atomic<int> bad{0};
  const uint64_t M = c->M;
  parallel_for(D, [&, snp_idx, M](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i)
      if (snp_idx[i] >= M || (i && snp_idx[i] <= snp_idx[i - 1])) bad.store(1);
  });
  if (bad.load()) return fail(c, GPCA_ERR_INVALID, "snp_idx must be strictly increasing and < num_snps");
  GPCA_TRY(ensure_counts(c));
  c->pca_idx.assign(snp_idx, snp_idx + D);
  c->h_mean.assign(mean, mean + D);
  c->h_sd.assign(sd, sd + D);
  return build_pca_set(c, D);
}

extern "C" int gpca_set_pca_snps_mask(gpca_ctx* c, const uint8_t* keep, const float* mean_all, const float* sd_all,
                                      uint64_t* n_pca_out) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->raw.p) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded (or it was released)");
  if (!keep || !mean_all || !sd_all) return fail(c, GPCA_ERR_INVALID, "null argument");
  GPCA_TRY(ensure_counts(c));
  const uint64_t M = c->M;
  // two-pass compaction over host threads: count per chunk, then fill
  const uint64_t CH = 1u << 18;
  const uint64_t nch = (M + CH - 1) / CH;
  std::vector<uint64_t> cnt(nch + 1, 0);
  parallel_for(nch, [&, keep, M](uint64_t lo, uint64_t hi) {
    for (uint64_t q = lo; q < hi; ++q) {
      uint64_t n = 0;
      const uint64_t e = std::min(M, (q + 1) * CH);
      for (uint64_t j = q * CH; j < e; ++j) n += keep[j] != 0;
      cnt[q + 1] = n;
    }
  }, 1);
  for (uint64_t q = 0; q < nch; ++q) cnt[q + 1] += cnt[q];
  const uint64_t D = cnt[nch];
  if (n_pca_out) *n_pca_out = D;
  if (D == 0) return fail(c, GPCA_ERR_INVALID, "No SNPs passed all QC filters.");  // prepare.rs:1020
  c->pca_idx.resize(D);
  c->h_mean.resize(D);
  c->h_sd.resize(D);
  uint64_t* pi = c->pca_idx.data();
  float* pm = c->h_mean.data();
  float* ps = c->h_sd.data();
  parallel_for(nch, [&, keep, M, mean_all, sd_all, pi, pm, ps](uint64_t lo, uint64_t hi) {
    for (uint64_t q = lo; q < hi; ++q) {
      uint64_t o = cnt[q];
      const uint64_t e = std::min(M, (q + 1) * CH);
      for (uint64_t j = q * CH; j < e; ++j)
        if (keep[j]) {
          pi[o] = j;
          pm[o] = mean_all[j];
          ps[o] = sd_all[j];
          ++o;
        }
    }
  }, 1);
  return build_pca_set(c, D);
}

// ---- one-call pipelined ingest ------------------------------------------------------------------------------------
// gpca_load_bed + (gpca_snp_qc | gpca_vcf_maf_filter) + gpca_set_pca_snps_mask as ONE streaming pass: while chunk q
// crosses PCIe, chunk q-1 is repitched and counted on the device, its 16-byte count records come back, the QC ladder
// runs on host threads, and the rows that pass are recoded into the resident SNP-major matrix.  The host -> device copy
// of the payload is the critical path; everything else hides behind it.  Replaces, for the data-preparation stage,
// MicroarrayDataPreparer::prepare_data_for_eigen_snp_pca (src/prepare.rs:995-1098) / the VCF read + MAF filter
// (src/vcf.rs:227-266, src/main.rs:176-212).
// payload source: host memory (host_payload) or, when fd >= 0, a file read chunk by chunk into pinned buffers
static int ingest_core(gpca_ctx* c, const uint8_t* host_payload, int fd, uint64_t file_offset, uint64_t n_in,
                       uint64_t n_snps, const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg,
                       double vcf_maf_threshold, uint8_t* keep_out, float* mean_out, float* sd_out,
                       uint8_t* fail_code_out, uint64_t* n_pca_out) {
  CHECK_CTX(c);
  const auto t_entry = std::chrono::steady_clock::now();
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!host_payload && fd < 0 && n_snps) return fail(c, GPCA_ERR_INVALID, "null payload");
  if (n_in == 0) return fail(c, GPCA_ERR_INVALID, "no samples");
  const uint64_t N = keep_samples ? n_keep : n_in;
  if (N == 0) return fail(c, GPCA_ERR_INVALID, "No samples passed QC.");  // prepare.rs:1010
  if (keep_samples)
    for (uint64_t i = 0; i < n_keep; ++i)
      if (keep_samples[i] < 0 || (uint64_t)keep_samples[i] >= n_in || (i && keep_samples[i] <= keep_samples[i - 1]))
        return fail(c, GPCA_ERR_INVALID, "keep_samples must be increasing indices into the FAM order");
  if (n_snps == 0) return fail(c, GPCA_ERR_INVALID, "No SNPs passed all QC filters.");
  c->vcf_mode = false;
  GPCA_TRY(alloc_raw(c, N, n_snps));
  const uint64_t M = n_snps;
  DevBuf<int64_t> d_keep;
  if (keep_samples) {
    GPCA_CUDA_TRY(c, d_keep.alloc(n_keep));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_keep.p, keep_samples, n_keep * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  }
  // outputs the caller did not ask for still exist internally
  std::vector<uint8_t> keep_tmp;
  std::vector<float> mean_tmp, sd_tmp;
  if (!keep_out) { keep_tmp.resize(M); keep_out = keep_tmp.data(); }
  if (!mean_out) { mean_tmp.resize(M); mean_out = mean_tmp.data(); }
  if (!sd_out) { sd_tmp.resize(M); sd_out = sd_tmp.data(); }

  const size_t in_pitch = (n_in + 3) / 4;
  uint64_t rows_per_chunk = std::max<uint64_t>(1, (128ull << 20) / std::max<size_t>(in_pitch, 1));
  if (const char* e = getenv("GPCA_INGEST_CHUNK_ROWS")) rows_per_chunk = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  rows_per_chunk = std::min<uint64_t>(rows_per_chunk, M);
  const uint64_t n_chunks = (M + rows_per_chunk - 1) / rows_per_chunk;

  // device-side destinations (worst case: every SNP passes)
  GPCA_CUDA_TRY(c, c->d_cnt.alloc(M));
  GPCA_CUDA_TRY(c, c->d_mean.alloc(M));
  GPCA_CUDA_TRY(c, c->d_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_inv_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_mu_inv_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_idx.alloc(M));
  c->Gs.cols = N;
  c->Gs.pitch = round_up((N + 3) / 4, 128);
  GPCA_CUDA_TRY(c, c->gs_store.alloc(c->Gs.pitch * M));
  c->Gs.p = c->gs_store.p;
  // Sample-major copy built behind the transfer: when memory allows (staging copy, Gs and a Gt whose row pitch covers
  // all M SNPs at once) every chunk's complete 512-row tiles of Gs are transposed as soon as they exist; only the tail
  // is left for the end of the call (the whole-matrix transpose was a 9 ms tail at 2,504 x 10M).  The pitch is then
  // the one for M columns whatever the QC keeps; the matrix is zeroed once, columns past D stay zero.
  bool incr_t = false;
  uint64_t t_done = 0;
  {
    const size_t pitch_max = round_up((M + 3) / 4, 128);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (!getenv("GPCA_DEBUG_NO_INCR_TRANSPOSE") &&
        (c->gt_store.n >= pitch_max * N || free_b > pitch_max * N + (16ull << 30))) {
      GPCA_CUDA_TRY(c, c->gt_store.alloc(pitch_max * N));
      c->Gt.p = c->gt_store.p;
      c->Gt.pitch = pitch_max;
      c->Gt.rows = N;
      GPCA_CUDA_TRY(c, cudaMemsetAsync(c->Gt.p, 0, pitch_max * N, c->stream));
      incr_t = true;
    }
  }
  auto transpose_rows = [&](uint64_t r_begin, uint64_t r_end) -> int {   // Gs rows [r_begin, r_end) -> Gt columns
    if (r_end <= r_begin) return GPCA_OK;
    PackedMat sv = c->Gs, dv = c->Gt;
    sv.p = c->Gs.p + r_begin * c->Gs.pitch;
    sv.rows = r_end - r_begin;
    dv.p = c->Gt.p + r_begin / 4;          // r_begin is a multiple of 512: a multiple of 128 bytes
    return launch_transpose(c, sv, dv);
  };
  if (c->h_cnt_cap < M) {
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    c->h_cnt = nullptr;
    c->h_cnt_cap = 0;
    GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_cnt, M * sizeof(uint4)));
    c->h_cnt_cap = M;
  }
  // pinned staging for the per-chunk compacted vectors: idx (8) + mean, sd, 1/sd, mean/sd (4 x 4) bytes per SNP, x2
  const size_t up_bytes = rows_per_chunk * 24;
  if (c->h_up_cap < 2 * up_bytes) {
    if (c->h_up) cudaFreeHost(c->h_up);
  for (int i = 0; i < 2; ++i)
    if (c->h_rd[i]) cudaFreeHost(c->h_rd[i]);
    c->h_up = nullptr;
    c->h_up_cap = 0;
    GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_up, 2 * up_bytes));
    c->h_up_cap = 2 * up_bytes;
  }
  if (fd >= 0) {   // two pinned read buffers, kept across calls
    const size_t need = rows_per_chunk * in_pitch;
    if (c->h_rd_cap < need) {
      for (int i = 0; i < 2; ++i) {
        if (c->h_rd[i]) cudaFreeHost(c->h_rd[i]);
        c->h_rd[i] = nullptr;
      }
      c->h_rd_cap = 0;
      for (int i = 0; i < 2; ++i) GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_rd[i], need));
      c->h_rd_cap = need;
    }
  }
  DevBuf<uint8_t>* stage = c->ingest_stage;   // kept across calls (no per-call cudaMalloc / cudaFree)
  // The payload crosses PCIe on its own stream: on the compute stream every chunk's transfer queued behind the
  // previous chunk's kernels (repitch, counts, recode, transpose tiles: ~0.45 ms per 128 MB chunk, 20 ms per 6.26 GB).
  if (!c->copy_stream) GPCA_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  const bool use_copy_stream = !getenv("GPCA_DEBUG_NO_COPY_STREAM");
  cudaEvent_t stage_free[2] = {nullptr, nullptr}, up_free[2] = {nullptr, nullptr}, h2d_done[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> cnt_ready(n_chunks, nullptr);
  auto cleanup = [&]() {
    cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
      if (stage_free[i]) cudaEventDestroy(stage_free[i]);
      if (up_free[i]) cudaEventDestroy(up_free[i]);
      if (h2d_done[i]) cudaEventDestroy(h2d_done[i]);
    }
    for (auto e : cnt_ready)
      if (e) cudaEventDestroy(e);
  };
  for (int i = 0; i < 2; ++i) {
    GPCA_CUDA_TRY(c, stage[i].alloc(rows_per_chunk * in_pitch + 16));
    cudaEventCreateWithFlags(&stage_free[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&up_free[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming);
  }
  for (auto& e : cnt_ready) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);

  if (c->h_counts.size() < M * 4) c->h_counts.resize(M * 4);
  if (c->pca_idx.size() < M) c->pca_idx.resize(M);
  if (c->h_mean.size() < M) c->h_mean.resize(M);
  if (c->h_sd.size() < M) c->h_sd.resize(M);
  if (c->h_inv.size() < M) c->h_inv.resize(M);
  if (c->h_muinv.size() < M) c->h_muinv.resize(M);
  const uint32_t pad = (uint32_t)(c->raw_pitch * 4 - N);
  const uint32_t n32 = (uint32_t)N;
  uint64_t D = 0, nmiss_total = 0;
  float inv_max = 0.f;
  int rc = GPCA_OK;

  const bool trace = getenv("GPCA_TRACE") != nullptr;
  double t_wait_cnt = 0, t_conv = 0, t_qc = 0, t_compact = 0, t_upload = 0, t_stage_wait = 0, t_enqueue = 0;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
  };
  const auto t_begin = now();
  // host half of the pipeline for chunk q (its count records have been requested on the stream already)
  auto process = [&](uint64_t q) -> int {
    const uint64_t r0 = q * rows_per_chunk, nr = std::min<uint64_t>(rows_per_chunk, M - r0);
    auto tp = now();
    GPCA_CUDA_TRY(c, cudaEventSynchronize(cnt_ready[q]));
    t_wait_cnt += ms_since(tp); tp = now();
    uint32_t* hc = c->h_counts.data();
    const uint4* hp = c->h_cnt;
    parallel_for(nr, [=](uint64_t lo, uint64_t hi) {
      for (uint64_t t = lo; t < hi; ++t) {
        const uint64_t j = r0 + t;
        const uint32_t miss = hp[j].x - pad, het = hp[j].y, d0 = hp[j].z;
        const uint32_t nv = n32 - miss;
        hc[4 * j + 0] = nv;
        hc[4 * j + 1] = d0;
        hc[4 * j + 2] = het;
        hc[4 * j + 3] = nv - d0 - het;
      }
    });
    t_conv += ms_since(tp); tp = now();
    if (cfg)
      host_snp_qc(N, nr, hc + 4 * r0, *cfg, keep_out + r0, mean_out + r0, sd_out + r0,
                  fail_code_out ? fail_code_out + r0 : nullptr);
    else
      host_vcf_maf(N, nr, hc + 4 * r0, vcf_maf_threshold, keep_out + r0, mean_out + r0, sd_out + r0);
    t_qc += ms_since(tp); tp = now();
    // compaction of the chunk (serial: a few hundred thousand SNPs) straight into the pinned upload buffer
    const int ub = (int)(q & 1);
    GPCA_CUDA_TRY(c, cudaEventSynchronize(up_free[ub]));
    uint8_t* ubase = c->h_up + (size_t)ub * up_bytes;
    uint64_t* u_idx = reinterpret_cast<uint64_t*>(ubase);
    float* u_mean = reinterpret_cast<float*>(ubase + rows_per_chunk * 8);
    float* u_sd = u_mean + rows_per_chunk;
    float* u_inv = u_sd + rows_per_chunk;
    float* u_mu = u_inv + rows_per_chunk;
    // Compaction on all host threads: the kept SNPs of fixed 8,192-row blocks are counted, the block offsets are a
    // short serial prefix sum, and every block then writes its survivors at its offset (serially this loop was 1.3 ms
    // per 128 MB chunk -- with it the host half of the pipeline took longer than the chunk's 2.4 ms on the bus).
    constexpr uint64_t CB = 8192;
    const uint64_t ncb = (nr + CB - 1) / CB;
    std::vector<uint64_t> cb_off(ncb + 1, 0), cb_miss(ncb, 0);
    std::vector<float> cb_inv(ncb, 0.f);
    parallel_for(ncb, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t b = lo; b < hi; ++b) {
        uint64_t cnt = 0;
        const uint64_t t1 = std::min<uint64_t>(nr, (b + 1) * CB);
        for (uint64_t t = b * CB; t < t1; ++t) cnt += keep_out[r0 + t] ? 1 : 0;
        cb_off[b + 1] = cnt;
      }
    }, 1);
    for (uint64_t b = 0; b < ncb; ++b) cb_off[b + 1] += cb_off[b];
    const uint64_t kept = cb_off[ncb];
    const uint64_t D0 = D;
    parallel_for(ncb, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t b = lo; b < hi; ++b) {
        uint64_t w = cb_off[b], nm = 0;
        float imax = 0.f;
        const uint64_t t1 = std::min<uint64_t>(nr, (b + 1) * CB);
        for (uint64_t t = b * CB; t < t1; ++t) {
          const uint64_t j = r0 + t;
          if (!keep_out[j]) continue;
          const float m = mean_out[j], sdv = sd_out[j];
          float inv = 0.f, mu = 0.f;
          if (!(std::fabs(sdv) < 1e-9f)) {   // same f32 expressions as prepare.rs:1948-1949
            inv = 1.0f / sdv;
            mu = m * inv;
          }
          u_idx[w] = j; u_mean[w] = m; u_sd[w] = sdv; u_inv[w] = inv; u_mu[w] = mu;
          c->pca_idx[D0 + w] = j; c->h_mean[D0 + w] = m; c->h_sd[D0 + w] = sdv;
          c->h_inv[D0 + w] = inv; c->h_muinv[D0 + w] = mu;
          imax = std::max(imax, std::fabs(inv));
          nm += N - hc[4 * j];
          ++w;
        }
        cb_miss[b] = nm;
        cb_inv[b] = imax;
      }
    }, 1);
    for (uint64_t b = 0; b < ncb; ++b) {
      nmiss_total += cb_miss[b];
      inv_max = std::max(inv_max, cb_inv[b]);
    }
    t_compact += ms_since(tp); tp = now();
    if (kept) {
      auto up = [&](void* dst, const void* src, size_t bytes) {
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
      };
      GPCA_CUDA_TRY(c, up(c->d_idx.p + D, u_idx, kept * 8));
      GPCA_CUDA_TRY(c, up(c->d_mean.p + D, u_mean, kept * 4));
      GPCA_CUDA_TRY(c, up(c->d_sd.p + D, u_sd, kept * 4));
      GPCA_CUDA_TRY(c, up(c->d_inv_sd.p + D, u_inv, kept * 4));
      GPCA_CUDA_TRY(c, up(c->d_mu_inv_sd.p + D, u_mu, kept * 4));
      PackedMat view = c->Gs;
      view.p = c->Gs.p + D * c->Gs.pitch;
      view.rows = kept;
      GPCA_TRY(launch_build_gs(c, c->raw.p, c->raw_pitch, c->d_idx.p + D, view));
    }
    GPCA_CUDA_TRY(c, cudaEventRecord(up_free[ub], c->stream));
    t_upload += ms_since(tp);
    D += kept;
    if (incr_t) {
      const uint64_t t_end = D & ~511ull;
      GPCA_TRY(transpose_rows(t_done, t_end));
      if (t_end > t_done) t_done = t_end;
    }
    return GPCA_OK;
  };

  for (uint64_t q = 0; q < n_chunks && rc == GPCA_OK; ++q) {
    const int buf = (int)(q & 1);
    const uint64_t r0 = q * rows_per_chunk, nr = std::min<uint64_t>(rows_per_chunk, M - r0);
    auto tq = now();
    cudaError_t e = cudaEventSynchronize(stage_free[buf]);   // (also: the copy out of pinned read buffer `buf` is done)
    t_stage_wait += ms_since(tq); tq = now();
    const uint8_t* src = host_payload ? host_payload + r0 * in_pitch : nullptr;
    if (e == cudaSuccess && fd >= 0) {
      // the read of chunk q overlaps with the transfer and the device work of chunk q-1 (already enqueued)
      size_t got = 0;
      const size_t want = nr * in_pitch;
      while (got < want) {
        const ssize_t r = pread(fd, c->h_rd[buf] + got, want - got, (off_t)(file_offset + r0 * in_pitch + got));
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got != want) {
        rc = fail(c, GPCA_ERR_INVALID, "ingest: short read from the .bed file");
        break;
      }
      src = c->h_rd[buf];
    }
    // staging buffer `buf` is free once the repitch of the chunk that used it last has run (compute stream)
    if (use_copy_stream) {
      if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, stage_free[buf], 0);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(stage[buf].p, src, nr * in_pitch, cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(h2d_done[buf], c->copy_stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, h2d_done[buf], 0);
    } else if (e == cudaSuccess) {
      e = cudaMemcpyAsync(stage[buf].p, src, nr * in_pitch, cudaMemcpyHostToDevice, c->stream);
    }
    if (e != cudaSuccess) {
      rc = fail(c, GPCA_ERR_CUDA, std::string("ingest H2D: ") + cudaGetErrorString(e));
      break;
    }
    rc = launch_repitch_gather(c, stage[buf].p, in_pitch, n_in, keep_samples ? d_keep.p : nullptr, N, nr,
                               c->raw.p + r0 * c->raw_pitch, c->raw_pitch);
    cudaEventRecord(stage_free[buf], c->stream);
    if (rc == GPCA_OK) rc = launch_bed_counts(c, c->raw.p + r0 * c->raw_pitch, c->raw_pitch, nr, c->d_cnt.p + r0);
    if (rc == GPCA_OK &&
        cudaMemcpyAsync(c->h_cnt + r0, c->d_cnt.p + r0, nr * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = fail(c, GPCA_ERR_CUDA, "ingest: count read-back failed");
    cudaEventRecord(cnt_ready[q], c->stream);
    t_enqueue += ms_since(tq);
    if (rc == GPCA_OK && q > 0) rc = process(q - 1);      // overlaps with chunk q's transfer
  }
  if (rc == GPCA_OK) rc = process(n_chunks - 1);
  if (rc != GPCA_OK) {
    cudaStreamSynchronize(c->stream);
    cleanup();
    reset_loaded(c);
    return rc;
  }
  c->have_counts = true;
  if (n_pca_out) *n_pca_out = D;
  if (D == 0) {
    cudaStreamSynchronize(c->stream);
    cleanup();
    return fail(c, GPCA_ERR_INVALID, "No SNPs passed all QC filters.");  // prepare.rs:1020
  }
  // (the host vectors stay at their capacity; only the first D entries are meaningful)
  c->any_missing = nmiss_total > 0;
  c->inv_sd_max = inv_max;
  c->D = D;
  c->Gs.rows = D;
  c->Gt.rows = N;
  c->Gt.cols = D;
  const double t_loop = ms_since(t_begin);
  if (incr_t) {
    rc = transpose_rows(t_done, D);        // the tail (fewer than 512 rows past the last complete tile)
  } else {
    c->Gt.pitch = round_up((D + 3) / 4, 128);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (c->gt_store.n < c->Gt.pitch * N && free_b < c->Gt.pitch * N + (8ull << 30)) {
      cudaStreamSynchronize(c->stream);
      c->raw.release();
    }
    cudaError_t ea = c->gt_store.alloc(c->Gt.pitch * N);
    if (ea != cudaSuccess) {
      cleanup();
      return fail(c, GPCA_ERR_OOM, std::string("ingest: ") + cudaGetErrorString(ea));
    }
    c->Gt.p = c->gt_store.p;
    rc = launch_transpose(c, c->Gs, c->Gt);
  }
  cudaStreamSynchronize(c->stream);
  cleanup();
  if (trace)
    fprintf(stderr,
            "[gpca_ingest_bed] chunks %llu  loop %.1f ms  total %.1f ms | stage wait %.1f  enqueue %.1f  count wait %.1f  "
            "convert %.1f  qc %.1f  compact %.1f  upload+build %.1f  (since entry %.1f)\n",
            (unsigned long long)n_chunks, t_loop, ms_since(t_begin), t_stage_wait, t_enqueue, t_wait_cnt, t_conv, t_qc,
            t_compact, t_upload, ms_since(t_entry));
  return rc;
}

extern "C" int gpca_ingest_bed(gpca_ctx* c, const uint8_t* host_payload, uint64_t n_in, uint64_t n_snps,
                               const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg,
                               double vcf_maf_threshold, uint8_t* keep_out, float* mean_out, float* sd_out,
                               uint8_t* fail_code_out, uint64_t* n_pca_out) {
  return ingest_core(c, host_payload, -1, 0, n_in, n_snps, keep_samples, n_keep, cfg, vcf_maf_threshold, keep_out,
                     mean_out, sd_out, fail_code_out, n_pca_out);
}

// Same pass straight from a PLINK .bed file: chunks are read into pinned buffers (never the whole file in host memory)
// and the read of chunk q overlaps with the transfer / counting / QC of the chunks before it.
extern "C" int gpca_ingest_bed_file(gpca_ctx* c, const char* bed_path, uint64_t n_in, uint64_t n_snps,
                                    const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg,
                                    double vcf_maf_threshold, uint8_t* keep_out, float* mean_out, float* sd_out,
                                    uint8_t* fail_code_out, uint64_t* n_pca_out) {
  CHECK_CTX(c);
  if (!bed_path) return fail(c, GPCA_ERR_INVALID, "null path");
  const int fd = open(bed_path, O_RDONLY);
  if (fd < 0) return fail(c, GPCA_ERR_INVALID, std::string("Failed to open BED file '") + bed_path + "'");
  uint8_t magic[3] = {0, 0, 0};
  struct stat st;
  int rc = GPCA_OK;
  if (pread(fd, magic, 3, 0) != 3 || magic[0] != 0x6c || magic[1] != 0x1b || magic[2] != 0x01)
    rc = fail(c, GPCA_ERR_INVALID, "not a SNP-major PLINK .bed (magic 6c 1b 01 expected)");
  else if (fstat(fd, &st) != 0 || (uint64_t)st.st_size != 3 + ((n_in + 3) / 4) * n_snps)
    rc = fail(c, GPCA_ERR_INVALID, "BED size does not match BIM/FAM");
  else
    rc = ingest_core(c, nullptr, fd, 3, n_in, n_snps, keep_samples, n_keep, cfg, vcf_maf_threshold, keep_out, mean_out,
                     sd_out, fail_code_out, n_pca_out);
  close(fd);
  return rc;
}

// ---- benchmark input ------------------------------------------------------------------------------------------
extern "C" int gpca_synth_bed_device(gpca_ctx* c, uint8_t* dev_out, uint64_t n_samples, uint64_t n_snps,
                                     uint64_t snp_offset, uint64_t seed, uint32_t n_pops, double fst,
                                     double missing_rate) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!dev_out || n_samples == 0) return fail(c, GPCA_ERR_INVALID, "synth: null output or no samples");
  GPCA_TRY(launch_synth_bed(c, dev_out, n_samples, n_snps, snp_offset, seed, n_pops, (float)fst, (float)missing_rate));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return GPCA_OK;
}

// ---- accessor parity ---------------------------------------------------------------------------
extern "C" int gpca_get_standardized_block(gpca_ctx* c, const uint64_t* ids, uint64_t n_ids, const uint64_t* samp,
                                           uint64_t n_samp, float* host_out) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->D == 0) return fail(c, GPCA_ERR_INVALID, "gpca_set_pca_snps has not been called");
  if (n_ids == 0 || n_samp == 0) return GPCA_OK;  // prepare.rs:1848 -> empty array
  if (!ids || !host_out) return fail(c, GPCA_ERR_INVALID, "null argument");
  for (uint64_t i = 0; i < n_ids; ++i)
    if (ids[i] >= c->D) return fail(c, GPCA_ERR_INVALID, "PcaSnpId out of range");
  if (samp)
    for (uint64_t j = 0; j < n_samp; ++j)
      if (samp[j] >= c->N) return fail(c, GPCA_ERR_INVALID, "QcSampleId out of range");
  if (!samp && n_samp != c->N) return fail(c, GPCA_ERR_INVALID, "sample list null but n_samp != num_samples");
  DevBuf<uint64_t> d_ids, d_samp;
  DevBuf<float> d_out;
  DevBuf<int> d_flag;
  GPCA_CUDA_TRY(c, d_ids.alloc(n_ids));
  GPCA_CUDA_TRY(c, d_out.alloc(n_ids * n_samp));
  GPCA_CUDA_TRY(c, d_flag.alloc(1));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_ids.p, ids, n_ids * 8, cudaMemcpyHostToDevice, c->stream));
  if (samp) {
    GPCA_CUDA_TRY(c, d_samp.alloc(n_samp));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_samp.p, samp, n_samp * 8, cudaMemcpyHostToDevice, c->stream));
  }
  GPCA_CUDA_TRY(c, cudaMemsetAsync(d_flag.p, 0, sizeof(int), c->stream));
  GPCA_TRY(launch_std_block(c, c->Gs, c->d_mean.p, c->d_sd.p, d_ids.p, n_ids, samp ? d_samp.p : nullptr, n_samp,
                            d_out.p, d_flag.p));
  int flag = 0;
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(host_out, d_out.p, n_ids * n_samp * 4, cudaMemcpyDeviceToHost, c->stream));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (flag) return fail(c, GPCA_ERR_MISSING, "Unexpected missing genotype in SnpBlockData (prepare.rs:1910)");
  return GPCA_OK;
}

// ---- sketch passes ---------------------------------------------------------------------------------
int timed_sketch(gpca_ctx* c, const SketchProblem& p) {
  cudaEvent_t e0, e1;
  GPCA_CUDA_TRY(c, cudaEventCreate(&e0));
  GPCA_CUDA_TRY(c, cudaEventCreate(&e1));
  GPCA_CUDA_TRY(c, cudaEventRecord(e0, c->stream));
  const int rc = launch_sketch(c, p);
  GPCA_CUDA_TRY(c, cudaEventRecord(e1, c->stream));
  c->pending_events.push_back({e0, e1});
  c->sk_bytes += (double)p.G.rows * (double)((p.G.cols + 3) / 4);
  c->sk_passes += 1;
  return rc;
}

int timed_sketch_batch(gpca_ctx* c, const SketchBatch& sb) {
  cudaEvent_t e0, e1;
  GPCA_CUDA_TRY(c, cudaEventCreate(&e0));
  GPCA_CUDA_TRY(c, cudaEventCreate(&e1));
  GPCA_CUDA_TRY(c, cudaEventRecord(e0, c->stream));
  const int rc = launch_sketch_i8_batch(c, sb);
  GPCA_CUDA_TRY(c, cudaEventRecord(e1, c->stream));
  c->pending_events.push_back({e0, e1});
  c->sk_bytes += sb.bytes;
  c->sk_passes += 1;
  return rc;
}

int sketch_snp_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld_in, uint32_t ld_out,
                    bool emit_stats, bool out_pad) {
  SketchProblem p;
  p.out_pad = out_pad;
  p.emit_stats = emit_stats;   // (the statistic is the max-abs over the l logical columns: independent of the stride)
  p.G = c->Gs;
  p.Bin = dev_in;
  p.l = l;
  p.ld = ld_in;
  p.f = nullptr;
  p.e = nullptr;
  p.a = c->d_inv_sd.p;
  p.b = c->d_mu_inv_sd.p;
  p.out = dev_out;
  p.ldo = ld_out;
  return timed_sketch(c, p);
}

int sketch_sample_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld_in, uint32_t ld_out,
                       bool use_stats) {
  SketchProblem p;
  p.use_stats = use_stats;
  p.G = c->Gt;
  p.Bin = dev_in;
  p.l = l;
  p.ld = ld_in;
  p.f = c->d_inv_sd.p;
  p.e = c->d_mu_inv_sd.p;
  p.a = nullptr;
  p.b = nullptr;
  p.out = dev_out;
  p.ldo = ld_out;
  GPCA_TRY(timed_sketch(c, p));
  if (c->allreduce) {
    if (ld_out != l) return fail(c, GPCA_ERR_INVALID, "sharded sample-side sketch needs ld == l");
    if (c->allreduce(dev_out, c->N * (uint64_t)l, 0, (void*)c->stream, c->allreduce_user) != 0)
      return fail(c, GPCA_ERR_CUDA, "allreduce hook failed");
  }
  return GPCA_OK;
}

int sketch_sample_side_gaussian(gpca_ctx* c, float* dev_scratch, float* dev_out, uint32_t l, uint32_t ld_in,
                                uint32_t ld_out, uint64_t seed, uint32_t stream_id) {
  SketchProblem p;
  p.G = c->Gt;
  p.Bin = dev_scratch;
  p.l = l;
  p.ld = ld_in;
  p.f = c->d_inv_sd.p;
  p.e = c->d_mu_inv_sd.p;
  p.a = nullptr;
  p.b = nullptr;
  p.out = dev_out;
  p.ldo = ld_out;
  p.gen = true;
  p.gen_seed = seed;
  p.gen_stream = stream_id;
  p.gen_row0 = c->shard_offset;
  p.gen_amax = GPCA_NORMAL_ABS_MAX * c->inv_sd_max;
  GPCA_TRY(timed_sketch(c, p));
  if (c->allreduce) {
    if (ld_out != l) return fail(c, GPCA_ERR_INVALID, "sharded sample-side sketch needs ld == l");
    if (c->allreduce(dev_out, c->N * (uint64_t)l, 0, (void*)c->stream, c->allreduce_user) != 0)
      return fail(c, GPCA_ERR_CUDA, "allreduce hook failed");
  }
  return GPCA_OK;
}

extern "C" int gpca_sketch_snp_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->D == 0) return fail(c, GPCA_ERR_INVALID, "gpca_set_pca_snps has not been called");
  if (l == 0 || l > 64 || ld < l) return fail(c, GPCA_ERR_INVALID, "need 1 <= l <= 64 and ld >= l");
  return sketch_snp_side(c, dev_in, dev_out, l, ld, ld, false, false);
}

extern "C" int gpca_sketch_sample_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (c->D == 0) return fail(c, GPCA_ERR_INVALID, "gpca_set_pca_snps has not been called");
  if (l == 0 || l > 64 || ld < l) return fail(c, GPCA_ERR_INVALID, "need 1 <= l <= 64 and ld >= l");
  return sketch_sample_side(c, dev_in, dev_out, l, ld, ld, false);
}

extern "C" double gpca_sketch_kernel_ms(gpca_ctx* c) { return c ? c->sk_kernel_ms_last : 0.0; }

extern "C" int gpca_sketch_stats(gpca_ctx* c, double* ms_total, double* bytes_total, uint64_t* n_passes, int reset) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (auto& pr : c->pending_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) c->sk_ms += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  c->pending_events.clear();
  for (auto& pr : c->pending_kernel_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) c->sk_kernel_ms += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  c->pending_kernel_events.clear();
  if (ms_total) *ms_total = c->sk_ms;
  if (bytes_total) *bytes_total = c->sk_bytes;
  if (n_passes) *n_passes = c->sk_passes;
  c->sk_kernel_ms_last = c->sk_kernel_ms;
  if (reset) {
    c->sk_kernel_ms = 0;
    c->sk_ms = 0;
    c->sk_bytes = 0;
    c->sk_passes = 0;
  }
  return GPCA_OK;
}
