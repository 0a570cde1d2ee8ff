// api.cu -- the C ABI (include/gpca.h): context, ingest, statistics, sketch entry points.
// The PCA drivers (rfit, EigenSNP) live in drivers.cu.
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cmath>
#include <cstring>
#include <atomic>
#include <functional>
#include <new>

#include "host_qc.h"
#include "kernels.cuh"
#include "sketch_tc.cuh"
#include "dense_tc.cuh"
#include "driver_util.cuh"
#include "parallel_for.h"

#define CHECK_CTX(c) \
  if (!(c)) return GPCA_ERR_INVALID;


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
  if (const char* ht = getenv("GPCA_HOST_THREADS")) c->host_threads = (unsigned)std::max(0, atoi(ht));
  if (const char* st = getenv("GPCA_SKETCH_TIMING")) c->sk_timing = atoi(st) != 0;
  *out = c;
  return GPCA_OK;
}

extern "C" void gpca_destroy(gpca_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  gpca_comm_destroy(c);      // before the stream it issued its collectives on goes away
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
  for (int i = 0; i < gpca_ctx::INGEST_STAGES; ++i)
    if (c->h_rd[i]) cudaFreeHost(c->h_rd[i]);
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
extern "C" int gpca_last_sketch_engine(const gpca_ctx* c) { return c ? c->last_engine : -1; }
extern "C" int gpca_set_batch_blocks(gpca_ctx* c, int on) {
  CHECK_CTX(c);
  c->batch_blocks = on ? 1 : 0;
  return GPCA_OK;
}
extern "C" int gpca_set_host_threads(gpca_ctx* c, uint32_t n) {
  CHECK_CTX(c);
  if (c->host_threads != n) c->pool.reset();      // re-created at the new size on next use
  c->host_threads = n;
  return GPCA_OK;
}
// The calling thread (and every thread it creates afterwards: the context's host pool, the caller's own workers) is
// bound to the CPUs next to the context's GPU, as sysfs reports them for the PCI device.  Pinned payload buffers that
// are first touched after this call then live in that NUMA node's memory: with one process per GPU on a two-socket
// host, half of the ranks otherwise stream their payload across the socket interconnect.
extern "C" int gpca_bind_host_to_device(gpca_ctx* c) {
  CHECK_CTX(c);
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, c->device) != cudaSuccess) return 0;
  for (char* p = bus; *p; ++p) *p = (char)tolower((unsigned char)*p);
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return 0;
  char line[4096] = {0};
  const bool ok = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!ok) return 0;
  cpu_set_t allowed, want;
  CPU_ZERO(&allowed);
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return 0;
  int n = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int lo = 0, hi = 0;
    if (sscanf(tok, "%d-%d", &lo, &hi) == 2) {
    } else if (sscanf(tok, "%d", &lo) == 1) {
      hi = lo;
    } else {
      continue;
    }
    for (int cpu = lo; cpu <= hi && cpu < CPU_SETSIZE; ++cpu)
      if (CPU_ISSET(cpu, &allowed)) {
        CPU_SET(cpu, &want);
        ++n;
      }
  }
  if (n == 0 || n == CPU_COUNT(&allowed)) return n;      // nothing to narrow (one NUMA node, or no overlap)
  if (sched_setaffinity(0, sizeof want, &want) != 0) return 0;
  c->pool.reset();                                       // (re-created on the bound CPUs on next use)
  return n;
}
extern "C" int gpca_set_sketch_timing(gpca_ctx* c, int on) {
  CHECK_CTX(c);
  c->sk_timing = on != 0;
  return GPCA_OK;
}
extern "C" int gpca_set_memory_reserve(gpca_ctx* c, uint64_t bytes) {
  CHECK_CTX(c);
  c->mem_reserve = bytes;
  return GPCA_OK;
}
extern "C" int gpca_set_ingest_mask(gpca_ctx* c, const uint8_t* mask, uint64_t n_snps) {
  CHECK_CTX(c);
  if (!mask || n_snps == 0) {
    c->ingest_mask.clear();
    return GPCA_OK;
  }
  c->ingest_mask.assign(mask, mask + n_snps);
  return GPCA_OK;
}
extern "C" uint64_t gpca_collective_count(const gpca_ctx* c) { return c ? c->collectives : 0; }
extern "C" uint64_t gpca_resident_snp_rows(const gpca_ctx* c) { return c ? (c->gs_win_rows ? c->gs_res_rows : c->D) : 0; }
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

// working buffers of gpca_eigensnp that survive between calls (their CONTENT is rebuilt by every call)
static void release_eigensnp_buffers(gpca_ctx* c) {
  c->es_store.release();
  c->et_store.release();
  c->ets_store.release();
  c->ess_store.release();
  c->es_cn.release();
  c->es_pool.release();
}
static size_t eigensnp_buffer_bytes(const gpca_ctx* c) {
  return c->es_store.n + c->et_store.n + c->ets_store.n + c->ess_store.n + c->es_cn.n * sizeof(float) + c->es_pool.bytes();
}

static void reset_loaded(gpca_ctx* c, bool keep_eigensnp_buffers = false) {
  c->data_version++;
  c->have_counts = false;   // (the host vectors keep their storage: re-growing them would zero-fill hundreds of MB)
  c->D = 0;
  c->Gs = PackedMat();
  c->Gt = PackedMat();   // (the stores are kept: DevBuf::alloc reuses them when the next data set fits)
  c->any_missing = false;
  c->gs_res_rows = 0;
  c->gs_win_rows = 0;
  if (!keep_eigensnp_buffers) release_eigensnp_buffers(c);
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

static int alloc_mapped(gpca_ctx* c, void** host, void** dev, size_t bytes) {
  GPCA_CUDA_TRY(c, cudaHostAlloc(host, bytes, cudaHostAllocMapped));
  GPCA_CUDA_TRY(c, cudaHostGetDevicePointer(dev, *host, 0));
  return GPCA_OK;
}

static int ensure_count_buffer(gpca_ctx* c, uint64_t M) {
  if (c->h_cnt_cap >= M && c->h_cnt) return GPCA_OK;
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  c->h_cnt = nullptr;
  c->h_cnt_dev = nullptr;
  c->h_cnt_cap = 0;
  GPCA_TRY(alloc_mapped(c, (void**)&c->h_cnt, (void**)&c->h_cnt_dev, std::max<uint64_t>(M, 1) * sizeof(uint4)));
  c->h_cnt_cap = M;
  return GPCA_OK;
}

// ---- statistics ----------------------------------------------------------------------------
static int ensure_counts(gpca_ctx* c) {
  if (c->have_counts) return GPCA_OK;
  GPCA_HOST_POOL(c);
  if (!c->raw.p) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded (or it was released)");
  const uint64_t M = c->M;
  GPCA_CUDA_TRY(c, c->d_cnt.alloc(std::max<uint64_t>(M, 1)));
  GPCA_TRY(launch_bed_counts(c, c->raw.p, c->raw_pitch, M, c->d_cnt.p));
  GPCA_TRY(ensure_count_buffer(c, M));   // pinned landing buffer for the 16 B/SNP count records, kept across calls
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
  GPCA_HOST_POOL(c);
  GPCA_TRY(ensure_counts(c));
  host_snp_qc(c->N, c->M, c->h_counts.data(), *cfg, keep, mean, sd, fail_code);
  return GPCA_OK;
}

extern "C" int gpca_vcf_maf_filter(gpca_ctx* c, double maf_threshold, uint8_t* keep, float* mean, float* sd) {
  CHECK_CTX(c);
  if (!keep) return fail(c, GPCA_ERR_INVALID, "keep null");
  GPCA_HOST_POOL(c);
  GPCA_TRY(ensure_counts(c));
  host_vcf_maf(c->N, c->M, c->h_counts.data(), maf_threshold, keep, mean, sd);
  return GPCA_OK;
}

// ---- PCA SNP set -----------------------------------------------------------------------------
// c->pca_idx, c->h_mean, c->h_sd hold the new set: derive 1/sd, mean/sd, the missing-call flag and the device vectors
static int derive_pca_vectors(gpca_ctx* c, uint64_t D) {
  GPCA_HOST_POOL(c);
  const uint64_t* idx = c->pca_idx.data();
  const float* mean = c->h_mean.data();
  const float* sd = c->h_sd.data();
  c->h_inv.resize(std::max<size_t>(c->h_inv.size(), D));
  c->h_muinv.resize(std::max<size_t>(c->h_muinv.size(), D));
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
  return GPCA_OK;
}

static int build_pca_set(gpca_ctx* c, uint64_t D) {
  // c->pca_idx, c->h_mean, c->h_sd are filled; derive the device-side vectors and the resident copies
  GPCA_TRY(derive_pca_vectors(c, D));
  c->data_version++;
  c->D = D;
  c->Gs.rows = D;
  c->Gs.cols = c->N;
  c->Gs.pitch = round_up((c->N + 3) / 4, 128);
  c->Gt.rows = c->N;
  c->Gt.cols = D;
  c->Gt.pitch = round_up((D + 3) / 4, 128);
  c->gs_res_rows = D;
  c->gs_win_rows = 0;
  GPCA_CUDA_TRY(c, c->gs_store.alloc(c->Gs.pitch * D));
  c->Gs.p = c->gs_store.p;
  GPCA_TRY(launch_build_gs(c, c->raw.p, c->raw_pitch, c->d_idx.p, c->Gs));
  // the PLINK-coded staging copy is only needed again for a different SNP selection; under memory pressure
  // (three copies would not fit comfortably) it is released before the transposed copy is allocated -- a later
  // gpca_set_pca_snps then narrows the resident set instead (reselect_resident)
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

// A new PCA SNP set when no staging copy of the payload exists (after the streaming ingest, or when build_pca_set
// released it): the set can only be narrowed.  The kept rows of the SNP-major matrix move up in place (chunk by chunk
// through a temporary: a row's source is never above its destination's chunk), the vectors are re-derived and the
// sample-major matrix is transposed again.  This is the call the EigenSNP workflow makes after mapping the QC'd SNPs
// to LD blocks (src/prepare.rs:1424-1563 drops SNPs outside every block).
static int reselect_resident(gpca_ctx* c, const uint64_t* snp_idx, uint64_t Dn, const float* mean, const float* sd) {
  if (c->D == 0 || !c->Gs.p) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded");
  if (c->gs_win_rows)
    return fail(c, GPCA_ERR_INVALID,
                "the SNP-major matrix is only partly resident: pass the SNP selection to the ingest (gpca_set_ingest_mask) "
                "instead of narrowing it afterwards");
  const uint64_t Do = c->D;
  std::vector<int64_t> pos(Dn);
  {
    uint64_t j = 0;
    for (uint64_t i = 0; i < Dn; ++i) {
      while (j < Do && c->pca_idx[j] < snp_idx[i]) ++j;
      if (j >= Do || c->pca_idx[j] != snp_idx[i])
        return fail(c, GPCA_ERR_INVALID,
                    "after a streaming ingest gpca_set_pca_snps can only narrow the resident SNP set (SNP not resident)");
      pos[i] = (int64_t)j++;
    }
  }
  const size_t pitch = c->Gs.pitch;
  const uint64_t chunk_rows = std::max<uint64_t>(512, (256ull << 20) / pitch);
  DevBuf<uint8_t> tmp;
  DevBuf<int64_t> d_pos;
  GPCA_CUDA_TRY(c, tmp.alloc(std::min<uint64_t>(chunk_rows, Dn) * pitch));
  GPCA_CUDA_TRY(c, d_pos.alloc(Dn));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(d_pos.p, pos.data(), Dn * 8, cudaMemcpyHostToDevice, c->stream));
  for (uint64_t n0 = 0; n0 < Dn; n0 += chunk_rows) {
    const uint64_t nr = std::min<uint64_t>(chunk_rows, Dn - n0);
    PackedMat dst = c->Gs;
    dst.p = tmp.p;
    dst.rows = nr;
    GPCA_TRY(launch_gather_rows(c, c->Gs, d_pos.p + n0, dst));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(c->Gs.p + n0 * pitch, tmp.p, nr * pitch, cudaMemcpyDeviceToDevice, c->stream));
  }
  const uint64_t old_written = std::min<uint64_t>(c->Gt.pitch, round_up(Do, 512) / 4);
  c->pca_idx.assign(snp_idx, snp_idx + Dn);
  c->h_mean.assign(mean, mean + Dn);
  c->h_sd.assign(sd, sd + Dn);
  GPCA_TRY(derive_pca_vectors(c, Dn));
  c->data_version++;
  c->D = Dn;
  c->Gs.rows = Dn;
  c->Gt.cols = Dn;
  c->gs_res_rows = Dn;
  GPCA_TRY(launch_transpose(c, c->Gs, c->Gt));
  const uint64_t new_written = std::min<uint64_t>(c->Gt.pitch, round_up(Dn, 512) / 4);
  if (new_written < old_written)
    GPCA_CUDA_TRY(c, cudaMemset2DAsync(c->Gt.p + new_written, c->Gt.pitch, 0, old_written - new_written, c->N, c->stream));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  // the EigenSNP copies cached in the context describe the old set
  c->es_store.release();
  c->et_store.release();
  c->ets_store.release();
  c->ess_store.release();
  c->es_cn.release();
  c->es_pool.release();
  return GPCA_OK;
}

extern "C" int gpca_set_pca_snps(gpca_ctx* c, const uint64_t* snp_idx, uint64_t D, const float* mean,
                                 const float* sd) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  GPCA_HOST_POOL(c);
  if (!c->raw.p && c->D == 0) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded");
  if (D == 0) return fail(c, GPCA_ERR_INVALID, "No SNPs passed all QC filters.");  // prepare.rs:1020
  if (!snp_idx || !mean || !sd) return fail(c, GPCA_ERR_INVALID, "null argument");
  std::atomic<int> bad{0};
  const uint64_t M = c->M;
  parallel_for(D, [&, snp_idx, M](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i)
      if (snp_idx[i] >= M || (i && snp_idx[i] <= snp_idx[i - 1])) bad.store(1);
  });
  if (bad.load()) return fail(c, GPCA_ERR_INVALID, "snp_idx must be strictly increasing and < num_snps");
  if (!c->raw.p) return reselect_resident(c, snp_idx, D, mean, sd);
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
  GPCA_HOST_POOL(c);
  if (!c->raw.p && c->D == 0) return fail(c, GPCA_ERR_INVALID, "no genotype payload loaded");
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
  std::vector<uint64_t> sel_idx;
  std::vector<float> sel_mean, sel_sd;
  const bool narrow = !c->raw.p;      // no staging copy: narrow the resident set (reselect_resident)
  if (narrow) {
    sel_idx.resize(D);
    sel_mean.resize(D);
    sel_sd.resize(D);
  } else {
    c->pca_idx.resize(D);
    c->h_mean.resize(D);
    c->h_sd.resize(D);
  }
  uint64_t* pi = narrow ? sel_idx.data() : c->pca_idx.data();
  float* pm = narrow ? sel_mean.data() : c->h_mean.data();
  float* ps = narrow ? sel_sd.data() : c->h_sd.data();
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
  if (narrow) return reselect_resident(c, pi, D, pm, ps);
  return build_pca_set(c, D);
}

// ---- one-call pipelined ingest ------------------------------------------------------------------------------------
// gpca_load_bed + (gpca_snp_qc | gpca_vcf_maf_filter) + gpca_set_pca_snps_mask as ONE streaming pass.  The host ->
// device copy of the payload is the critical path and runs back to back on its own stream, up to INGEST_LAG chunks
// ahead of the host; behind it, per chunk:
//   device: counts straight from the staged rows (file pitch, unaligned) -> 16-byte records written into MAPPED pinned
//           host memory by the kernel (no copy in the transfer queue)
//   host:   count records -> QC ladder on the pool threads -> compaction of the kept SNPs' index / mean / sd / 1/sd /
//           mean/sd into mapped pinned memory
//   device: a small kernel fetches those vectors, the kept rows are recoded from the staging buffer into the SNP-major
//           resident matrix, and every complete 512-row tile of it is transposed into the sample-major one.
// No full-size staging copy exists: at most three copies of a CHUNK (stage, optional sample-gathered copy) beside the
// two resident orientations.  When even those two do not fit (mem_reserve is left free for the drivers), the SNP-major
// matrix keeps only its first rows resident (gpca_ctx::gs_res_rows) and later rows pass through a ring window.
// Replaces, for the data-preparation stage, MicroarrayDataPreparer::prepare_data_for_eigen_snp_pca
// (src/prepare.rs:995-1098) / the VCF read + MAF filter (src/vcf.rs:227-266, src/main.rs:176-212).
// payload source: host memory (host_payload) or, when fd >= 0, a file read chunk by chunk into pinned buffers
// Gs rows [r_begin, r_end) (logical; both multiples of 512 except the very last end) -> Gt columns, following the
// physical placement of the rows (resident part, then the ring window)
static int transpose_gs_rows(gpca_ctx* c, uint64_t r_begin, uint64_t r_end) {
  const uint64_t res = c->gs_res_rows, win = c->gs_win_rows;
  uint64_t r = r_begin;
  while (r < r_end) {
    uint64_t seg_end, phys;
    if (r < res || win == 0) {
      seg_end = win ? std::min(r_end, res) : r_end;
      phys = r;
    } else {
      const uint64_t off = (r - res) % win;
      seg_end = std::min(r_end, r + (win - off));
      phys = res + off;
    }
    PackedMat sv = c->Gs, dv = c->Gt;
    sv.p = c->Gs.p + phys * c->Gs.pitch;
    sv.rows = seg_end - r;
    dv.p = c->Gt.p + r / 4;          // r is a multiple of 512: a multiple of 128 bytes
    GPCA_TRY(launch_transpose(c, sv, dv));
    r = seg_end;
  }
  return GPCA_OK;
}

static int ingest_core(gpca_ctx* c, const uint8_t* host_payload, int fd, uint64_t file_offset, uint64_t n_in,
                       uint64_t n_snps, const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg,
                       double vcf_maf_threshold, uint8_t* keep_out, float* mean_out, float* sd_out,
                       uint8_t* fail_code_out, uint64_t* n_pca_out) {
  CHECK_CTX(c);
  GPCA_HOST_POOL(c);
  constexpr int NS = gpca_ctx::INGEST_STAGES;
  constexpr uint64_t LAG = 2;       // payload chunks in the transfer queue ahead of the host (NS >= LAG + 2)
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
  const uint8_t* pre_mask = nullptr;
  if (!c->ingest_mask.empty()) {
    if (c->ingest_mask.size() != n_snps) return fail(c, GPCA_ERR_INVALID, "ingest mask length != n_snps");
    pre_mask = c->ingest_mask.data();
  }
  c->vcf_mode = false;
  // (a host that ingests and runs EigenSNP repeatedly keeps the driver's working buffers: freeing and re-allocating
  //  tens of GB per call costs more than the call's kernels; they are given up below if the matrices need the room)
  reset_loaded(c, true);
  c->raw.release();                 // (the three-call path's staging copy, if an earlier data set left one)
  c->raw_pitch = 0;
  c->N = N;
  c->M = n_snps;
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
  const size_t gather_pitch = round_up((N + 3) / 4, 16);
  uint64_t rows_per_chunk = std::max<uint64_t>(1, (128ull << 20) / std::max<size_t>(in_pitch, 1));
  if (const char* e = getenv("GPCA_INGEST_CHUNK_ROWS")) rows_per_chunk = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  rows_per_chunk = std::min<uint64_t>(rows_per_chunk, M);
  const uint64_t n_chunks = (M + rows_per_chunk - 1) / rows_per_chunk;

  // device-side destinations (worst case: every SNP passes)
  GPCA_CUDA_TRY(c, c->d_mean.alloc(M));
  GPCA_CUDA_TRY(c, c->d_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_inv_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_mu_inv_sd.alloc(M));
  GPCA_CUDA_TRY(c, c->d_idx.alloc(M));
  GPCA_TRY(ensure_count_buffer(c, M));
  DevBuf<uint8_t>* stage = c->ingest_stage;   // kept across calls (no per-call cudaMalloc / cudaFree)
  for (int i = 0; i < NS; ++i) {
    GPCA_CUDA_TRY(c, stage[i].alloc(rows_per_chunk * in_pitch + 64));
    if (keep_samples) GPCA_CUDA_TRY(c, c->ingest_gather[i].alloc(rows_per_chunk * gather_pitch + 64));
  }
  // pinned + mapped staging for the per-chunk compacted vectors: idx (8) + mean, sd, 1/sd, mean/sd (4 x 4) bytes per SNP
  const size_t up_bytes = round_up(rows_per_chunk * 24, 256);
  if (c->h_up_cap < NS * up_bytes) {
    if (c->h_up) cudaFreeHost(c->h_up);
    c->h_up = nullptr;
    c->h_up_dev = nullptr;
    c->h_up_cap = 0;
    GPCA_TRY(alloc_mapped(c, (void**)&c->h_up, (void**)&c->h_up_dev, NS * up_bytes));
    c->h_up_cap = NS * up_bytes;
  }
  if (fd >= 0) {   // pinned read buffers, kept across calls
    const size_t need = rows_per_chunk * in_pitch;
    if (c->h_rd_cap < need) {
      for (int i = 0; i < NS; ++i) {
        if (c->h_rd[i]) cudaFreeHost(c->h_rd[i]);
        c->h_rd[i] = nullptr;
      }
      c->h_rd_cap = 0;
      for (int i = 0; i < NS; ++i) GPCA_CUDA_TRY(c, cudaMallocHost((void**)&c->h_rd[i], need));
      c->h_rd_cap = need;
    }
  }

  // ---- layout of the two resident orientations under the memory budget ---------------------------------------------
  // Gt's pitch covers all M columns whatever the QC keeps (tiles are transposed while the count of survivors is still
  // open); columns past D are zeroed at the end.
  c->Gs.cols = N;
  c->Gs.pitch = round_up((N + 3) / 4, 128);
  c->Gt.pitch = round_up((M + 3) / 4, 128);
  c->Gt.rows = N;
  {
    const size_t gs_full = c->Gs.pitch * M, gt_full = c->Gt.pitch * N;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t avail = free_b + c->gs_store.n + c->gt_store.n;     // (stores of an earlier data set are reused or replaced)
    // the reserve is for the drivers' working buffers: what gpca_eigensnp already holds from an earlier call counts
    size_t reserve = c->mem_reserve > eigensnp_buffer_bytes(c) ? c->mem_reserve - eigensnp_buffer_bytes(c) : 0;
    if (gs_full + gt_full + reserve > avail && eigensnp_buffer_bytes(c) > 0) {
      // ... unless both orientations only fit without them (they are re-allocated inside the reserve when needed)
      GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      release_eigensnp_buffers(c);
      cudaMemGetInfo(&free_b, &total_b);
      avail = free_b + c->gs_store.n + c->gt_store.n;
      reserve = c->mem_reserve;
    }
    if (const char* e = getenv("GPCA_DEBUG_MEM_BUDGET")) avail = std::min<size_t>(avail, strtoull(e, nullptr, 10));
    uint64_t res = M, win = 0;
    if (gs_full + gt_full + reserve > avail) {
      win = round_up(2 * rows_per_chunk + 1024, 512);
      const uint64_t win_target = round_up(std::max<uint64_t>(1, (4ull << 30) / c->Gs.pitch), 512);   // ~4 GB of rows
      win = std::max(win, win_target);
      if (const char* e = getenv("GPCA_DEBUG_WINDOW_ROWS"))      // tests: a small window on a small matrix
        win = round_up(std::max<uint64_t>(strtoull(e, nullptr, 10), rows_per_chunk + 1024), 512);
      const size_t fixed = gt_full + reserve + win * c->Gs.pitch;
      res = fixed < avail ? ((avail - fixed) / c->Gs.pitch) & ~511ull : 0;
      if (res + win >= M) {      // the window alone covers the rest: everything is resident after all
        res = M;
        win = 0;
      }
      if (gt_full + std::min<size_t>(gs_full, win * c->Gs.pitch) > avail)
        return fail(c, GPCA_ERR_OOM, "ingest: the sample-major matrix does not fit on this device (shard the SNPs over more GPUs)");
    }
    const size_t gs_bytes = win ? (res + win) * c->Gs.pitch : gs_full;
    // release before re-allocating (the old and the new store together may not fit), and give back what a smaller
    // layout no longer needs (the reserve is for the drivers)
    if (c->gs_store.n < gs_bytes || c->gs_store.n > gs_bytes + (1ull << 30)) c->gs_store.release();
    if (c->gt_store.n < gt_full || c->gt_store.n > gt_full + (1ull << 30)) c->gt_store.release();
    GPCA_CUDA_TRY(c, c->gt_store.alloc(gt_full));
    GPCA_CUDA_TRY(c, c->gs_store.alloc(gs_bytes));
    c->Gs.p = c->gs_store.p;
    c->Gt.p = c->gt_store.p;
    c->gs_res_rows = res;
    c->gs_win_rows = win;
  }
  uint64_t t_done = 0;

  // The payload crosses PCIe on its own stream: on the compute stream every chunk's transfer queued behind the
  // previous chunk's kernels.
  if (!c->copy_stream) GPCA_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  cudaEvent_t stage_free[NS], up_free[NS], h2d_done[NS];
  for (int i = 0; i < NS; ++i) stage_free[i] = up_free[i] = h2d_done[i] = nullptr;
  std::vector<cudaEvent_t> cnt_ready(n_chunks, nullptr);
  auto cleanup = [&]() {
    cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < NS; ++i) {
      if (stage_free[i]) cudaEventDestroy(stage_free[i]);
      if (up_free[i]) cudaEventDestroy(up_free[i]);
      if (h2d_done[i]) cudaEventDestroy(h2d_done[i]);
    }
    for (auto e : cnt_ready)
      if (e) cudaEventDestroy(e);
  };
  for (int i = 0; i < NS; ++i) {
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
  // the rows of chunk q as the device sees them (staged file rows, or their sample-gathered copy)
  auto chunk_src = [&](uint64_t q, const uint8_t*& p, size_t& pitch) {
    const int sl = (int)(q % NS);
    if (keep_samples) {
      p = c->ingest_gather[sl].p;
      pitch = gather_pitch;
    } else {
      p = stage[sl].p;
      pitch = in_pitch;
    }
  };
  // device half, part 1: transfer of chunk q and its counts
  auto enqueue = [&](uint64_t q) -> int {
    const int sl = (int)(q % NS);
    const uint64_t r0 = q * rows_per_chunk, nr = std::min<uint64_t>(rows_per_chunk, M - r0);
    auto tq = now();
    const uint8_t* src = host_payload ? host_payload + r0 * in_pitch : nullptr;
    if (fd >= 0) {
      // the pinned read buffer is free once the copy that last used it has run; the read of chunk q then overlaps with
      // the transfers and the device work of the chunks before it
      GPCA_CUDA_TRY(c, cudaEventSynchronize(h2d_done[sl]));
      t_stage_wait += ms_since(tq); tq = now();
      size_t got = 0;
      const size_t want = nr * in_pitch;
      while (got < want) {
        const ssize_t r = pread(fd, c->h_rd[sl] + got, want - got, (off_t)(file_offset + r0 * in_pitch + got));
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got != want) return fail(c, GPCA_ERR_INVALID, "ingest: short read from the .bed file");
      src = c->h_rd[sl];
    }
    // staging buffer `sl` is free once the recode of the chunk that used it last has run (compute stream)
    GPCA_CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, stage_free[sl], 0));
    GPCA_CUDA_TRY(c, cudaMemcpyAsync(stage[sl].p, src, nr * in_pitch, cudaMemcpyHostToDevice, c->copy_stream));
    GPCA_CUDA_TRY(c, cudaEventRecord(h2d_done[sl], c->copy_stream));
    GPCA_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, h2d_done[sl], 0));
    if (keep_samples)
      GPCA_TRY(launch_repitch_gather(c, stage[sl].p, in_pitch, n_in, d_keep.p, N, nr, c->ingest_gather[sl].p, gather_pitch));
    const uint8_t* cs;
    size_t cp;
    chunk_src(q, cs, cp);
    GPCA_TRY(launch_chunk_counts(c, cs, cp, N, nr, c->h_cnt_dev + r0));
    GPCA_CUDA_TRY(c, cudaEventRecord(cnt_ready[q], c->stream));
    t_enqueue += ms_since(tq);
    return GPCA_OK;
  };
  // host half of the pipeline for chunk q, then device half part 2 (recode of the kept rows, transposes)
  auto process = [&](uint64_t q) -> int {
    const int sl = (int)(q % NS);
    const uint64_t r0 = q * rows_per_chunk, nr = std::min<uint64_t>(rows_per_chunk, M - r0);
    auto tp = now();
    GPCA_CUDA_TRY(c, cudaEventSynchronize(cnt_ready[q]));
    t_wait_cnt += ms_since(tp); tp = now();
    uint32_t* hc = c->h_counts.data();
    const uint4* hp = c->h_cnt;
    parallel_for(nr, [=](uint64_t lo, uint64_t hi) {
      for (uint64_t t = lo; t < hi; ++t) {
        const uint64_t j = r0 + t;
        const uint32_t miss = hp[j].x, het = hp[j].y, d0 = hp[j].z;
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
    if (pre_mask)      // rows outside the caller's pre-selection are dropped whatever the QC says (fail code 7)
      parallel_for(nr, [=](uint64_t lo, uint64_t hi) {
        for (uint64_t t = lo; t < hi; ++t)
          if (!pre_mask[r0 + t] && keep_out[r0 + t]) {
            keep_out[r0 + t] = 0;
            mean_out[r0 + t] = 0.f;
            sd_out[r0 + t] = 0.f;
            if (fail_code_out) fail_code_out[r0 + t] = 7;
          }
      });
    t_qc += ms_since(tp); tp = now();
    GPCA_CUDA_TRY(c, cudaEventSynchronize(up_free[sl]));
    const size_t uoff = (size_t)sl * up_bytes;
    uint8_t* ubase = c->h_up + uoff;
    uint64_t* u_idx = reinterpret_cast<uint64_t*>(ubase);
    float* u_mean = reinterpret_cast<float*>(ubase + rows_per_chunk * 8);
    float* u_sd = u_mean + rows_per_chunk;
    float* u_inv = u_sd + rows_per_chunk;
    float* u_mu = u_inv + rows_per_chunk;
    // Compaction on the pool threads: the kept SNPs of fixed 8,192-row blocks are counted, the block offsets are a
    // short serial prefix sum, and every block then writes its survivors at its offset.
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
      const uint8_t* ud = c->h_up_dev + uoff;
      const float* d_um = reinterpret_cast<const float*>(ud + rows_per_chunk * 8);
      GPCA_TRY(launch_fetch_kept(c, reinterpret_cast<const uint64_t*>(ud), d_um, d_um + rows_per_chunk,
                                 d_um + 2 * rows_per_chunk, d_um + 3 * rows_per_chunk, kept, c->d_idx.p + D,
                                 c->d_mean.p + D, c->d_sd.p + D, c->d_inv_sd.p + D, c->d_mu_inv_sd.p + D));
      const uint8_t* cs;
      size_t cp;
      chunk_src(q, cs, cp);
      GPCA_TRY(launch_build_gs_chunk(c, cs, cp, N, r0, c->d_idx.p + D, kept, c->Gs.p, c->Gs.pitch, D, c->gs_res_rows,
                                     c->gs_win_rows));
    }
    GPCA_CUDA_TRY(c, cudaEventRecord(up_free[sl], c->stream));
    GPCA_CUDA_TRY(c, cudaEventRecord(stage_free[sl], c->stream));
    t_upload += ms_since(tp);
    D += kept;
    const uint64_t t_end = D & ~511ull;
    if (t_end > t_done) {
      GPCA_TRY(transpose_gs_rows(c, t_done, t_end));
      t_done = t_end;
    }
    return GPCA_OK;
  };

  for (uint64_t q = 0; q < n_chunks && rc == GPCA_OK; ++q) {
    rc = enqueue(q);
    if (rc == GPCA_OK && q >= LAG) rc = process(q - LAG);
  }
  for (uint64_t q = (n_chunks > LAG ? n_chunks - LAG : 0); q < n_chunks && rc == GPCA_OK; ++q) rc = process(q);
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
  c->Gt.cols = D;
  if (c->gs_res_rows >= D) {       // everything ended up in the resident part
    c->gs_res_rows = D;
  }
  const double t_loop = ms_since(t_begin);
  rc = transpose_gs_rows(c, t_done, D);        // the tail (fewer than 512 rows past the last complete tile)
  {
    // columns of Gt past the last transposed tile were never written: zero them (a strided memset of the row tails)
    const size_t written = std::min<size_t>(c->Gt.pitch, round_up(D, 512) / 4);
    if (rc == GPCA_OK && written < c->Gt.pitch)
      GPCA_CUDA_TRY(c, cudaMemset2DAsync(c->Gt.p + written, c->Gt.pitch, 0, c->Gt.pitch - written, N, c->stream));
  }
  cudaStreamSynchronize(c->stream);
  cleanup();
  if (trace)
    fprintf(stderr,
            "[gpca_ingest_bed] chunks %llu  loop %.1f ms  total %.1f ms | stage wait %.1f  enqueue %.1f  count wait %.1f  "
            "convert %.1f  qc %.1f  compact %.1f  fetch+build %.1f  (since entry %.1f)  Gs resident rows %llu of %llu, window %llu\n",
            (unsigned long long)n_chunks, t_loop, ms_since(t_begin), t_stage_wait, t_enqueue, t_wait_cnt, t_conv, t_qc,
            t_compact, t_upload, ms_since(t_entry), (unsigned long long)c->gs_res_rows, (unsigned long long)D,
            (unsigned long long)c->gs_win_rows);
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
// and the read of chunk q overlaps with the transfer / counting / QC of the chunks before it.  The _rows form takes
// rows [first_row, first_row + n_rows) of the file: one shard of a multi-GPU run (outputs and the ingest mask are
// indexed by the row inside the range).
extern "C" int gpca_ingest_bed_file_rows(gpca_ctx* c, const char* bed_path, uint64_t n_in, uint64_t n_snps_in_file,
                                         uint64_t first_row, uint64_t n_rows, const int64_t* keep_samples,
                                         uint64_t n_keep, const gpca_qc_cfg* cfg, double vcf_maf_threshold,
                                         uint8_t* keep_out, float* mean_out, float* sd_out, uint8_t* fail_code_out,
                                         uint64_t* n_pca_out) {
  CHECK_CTX(c);
  if (!bed_path) return fail(c, GPCA_ERR_INVALID, "null path");
  if (first_row > n_snps_in_file || n_rows > n_snps_in_file - first_row)
    return fail(c, GPCA_ERR_INVALID, "row range outside the .bed file");
  const int fd = open(bed_path, O_RDONLY);
  if (fd < 0) return fail(c, GPCA_ERR_INVALID, std::string("Failed to open BED file '") + bed_path + "'");
  uint8_t magic[3] = {0, 0, 0};
  struct stat st;
  int rc = GPCA_OK;
  const uint64_t in_pitch = (n_in + 3) / 4;
  if (pread(fd, magic, 3, 0) != 3 || magic[0] != 0x6c || magic[1] != 0x1b || magic[2] != 0x01)
    rc = fail(c, GPCA_ERR_INVALID, "not a SNP-major PLINK .bed (magic 6c 1b 01 expected)");
  else if (fstat(fd, &st) != 0 || (uint64_t)st.st_size != 3 + in_pitch * n_snps_in_file)
    rc = fail(c, GPCA_ERR_INVALID, "BED size does not match BIM/FAM");
  else
    rc = ingest_core(c, nullptr, fd, 3 + first_row * in_pitch, n_in, n_rows, keep_samples, n_keep, cfg,
                     vcf_maf_threshold, keep_out, mean_out, sd_out, fail_code_out, n_pca_out);
  close(fd);
  return rc;
}

extern "C" int gpca_ingest_bed_file(gpca_ctx* c, const char* bed_path, uint64_t n_in, uint64_t n_snps,
                                    const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg,
                                    double vcf_maf_threshold, uint8_t* keep_out, float* mean_out, float* sd_out,
                                    uint8_t* fail_code_out, uint64_t* n_pca_out) {
  return gpca_ingest_bed_file_rows(c, bed_path, n_in, n_snps, 0, n_snps, keep_samples, n_keep, cfg, vcf_maf_threshold,
                                   keep_out, mean_out, sd_out, fail_code_out, n_pca_out);
}

// ---- pinned host buffers for large payloads ----------------------------------------------------------------------
// cudaMallocHost pins 4 KB pages one fault at a time (measured 2.4 GB/s: 36 s for the 87.5 GB payload of 500,000 x
// 700,000).  Here: anonymous memory advised to transparent huge pages, first-touched on the context's host threads,
// then registered with CUDA in one call.
extern "C" void* gpca_host_alloc(gpca_ctx* c, uint64_t bytes) {
  if (!c || bytes == 0) return nullptr;
  GPCA_HOST_POOL(c);
  const uint64_t HP = 2ull << 20;
  const uint64_t len = round_up(bytes, HP);
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) {
    c->set_error("gpca_host_alloc: mmap failed");
    return nullptr;
  }
  madvise(p, len, MADV_HUGEPAGE);
  uint8_t* b = static_cast<uint8_t*>(p);
  parallel_for(len / 4096, [b](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) b[i * 4096] = 0;
  }, 1u << 12);
  if (cudaSetDevice(c->device) != cudaSuccess || cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
    cudaGetLastError();
    munmap(p, len);
    c->set_error("gpca_host_alloc: cudaHostRegister failed");
    return nullptr;
  }
  return p;
}
extern "C" void gpca_host_free(gpca_ctx* c, void* p, uint64_t bytes) {
  if (!p) return;
  if (c) cudaSetDevice(c->device);
  cudaHostUnregister(p);
  munmap(p, round_up(bytes, 2ull << 20));
}

// ---- benchmark input ------------------------------------------------------------------------------------------
extern "C" int gpca_synth_bed_device(gpca_ctx* c, uint8_t* dev_out, uint64_t n_samples, uint64_t n_snps,
                                     uint64_t snp_offset, uint64_t seed, uint32_t n_pops, double fst,
                                     double missing_rate, double fst_grade) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!dev_out || n_samples == 0) return fail(c, GPCA_ERR_INVALID, "synth: null output or no samples");
  GPCA_TRY(launch_synth_bed(c, dev_out, n_samples, n_snps, snp_offset, seed, n_pops, (float)fst, (float)missing_rate,
                            (float)fst_grade));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return GPCA_OK;
}

// Measurement hook for K-a alone: the allele-count kernel of the ingest (chunk_counts_kernel) on a payload that is
// already on the device, `reps` launches between two CUDA events; *ms_out = mean time of one launch.
extern "C" int gpca_count_kernel_ms(gpca_ctx* c, const uint8_t* dev_payload, uint64_t n_samples, uint64_t n_snps,
                                    uint32_t reps, double* ms_out) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!dev_payload || !ms_out || n_samples == 0 || n_snps == 0 || reps == 0)
    return fail(c, GPCA_ERR_INVALID, "gpca_count_kernel_ms: bad argument");
  GPCA_CUDA_TRY(c, c->d_cnt.alloc(n_snps));
  const size_t pitch = (n_samples + 3) / 4;
  GPCA_TRY(launch_chunk_counts(c, dev_payload, pitch, n_samples, n_snps, c->d_cnt.p));      // warm-up
  cudaEvent_t e0, e1;
  GPCA_CUDA_TRY(c, cudaEventCreate(&e0));
  GPCA_CUDA_TRY(c, cudaEventCreate(&e1));
  cudaEventRecord(e0, c->stream);
  int rc = GPCA_OK;
  for (uint32_t r = 0; r < reps && rc == GPCA_OK; ++r)
    rc = launch_chunk_counts(c, dev_payload, pitch, n_samples, n_snps, c->d_cnt.p);
  cudaEventRecord(e1, c->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = (double)ms / reps;
  return rc;
}

// the same generator into HOST memory (chunks are generated on the device and copied back): the payload a host would
// have read from a .bed file, for shapes whose payload does not fit on the device next to the resident matrices
extern "C" int gpca_synth_bed_host(gpca_ctx* c, uint8_t* host_out, uint64_t n_samples, uint64_t n_snps,
                                   uint64_t snp_offset, uint64_t seed, uint32_t n_pops, double fst, double missing_rate,
                                   double fst_grade) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!host_out || n_samples == 0) return fail(c, GPCA_ERR_INVALID, "synth: null output or no samples");
  const uint64_t bps = (n_samples + 3) / 4;
  const uint64_t rows_per_chunk = std::max<uint64_t>(1, (1ull << 30) / bps);
  DevBuf<uint8_t> buf[2];
  cudaEvent_t done[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    GPCA_CUDA_TRY(c, buf[i].alloc(std::min(rows_per_chunk, n_snps) * bps));
    GPCA_CUDA_TRY(c, cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
  }
  if (!c->copy_stream) GPCA_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  cudaEvent_t gen_done = nullptr;
  GPCA_CUDA_TRY(c, cudaEventCreateWithFlags(&gen_done, cudaEventDisableTiming));
  int rc = GPCA_OK;
  uint64_t q = 0;
  for (uint64_t r0 = 0; r0 < n_snps && rc == GPCA_OK; r0 += rows_per_chunk, ++q) {
    const int b = (int)(q & 1);
    const uint64_t nr = std::min(rows_per_chunk, n_snps - r0);
    cudaStreamWaitEvent(c->stream, done[b], 0);      // the copy that last read this buffer has run
    rc = launch_synth_bed(c, buf[b].p, n_samples, nr, snp_offset + r0, seed, n_pops, (float)fst, (float)missing_rate,
                          (float)fst_grade);
    cudaEventRecord(gen_done, c->stream);
    cudaStreamWaitEvent(c->copy_stream, gen_done, 0);
    if (cudaMemcpyAsync(host_out + r0 * bps, buf[b].p, nr * bps, cudaMemcpyDeviceToHost, c->copy_stream) != cudaSuccess)
      rc = fail(c, GPCA_ERR_CUDA, "synth: copy to the host failed");
    cudaEventRecord(done[b], c->copy_stream);
  }
  cudaStreamSynchronize(c->copy_stream);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]);
  cudaEventDestroy(gen_done);
  return rc;
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
  GPCA_TRY(launch_std_block(c, c->Gs, c->gs_win_rows ? c->gs_res_rows : c->D, c->Gt, c->d_mean.p, c->d_sd.p, d_ids.p, n_ids, samp ? d_samp.p : nullptr, n_samp,
                            d_out.p, d_flag.p));
  int flag = 0;
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GPCA_CUDA_TRY(c, cudaMemcpyAsync(host_out, d_out.p, n_ids * n_samp * 4, cudaMemcpyDeviceToHost, c->stream));
  GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (flag) return fail(c, GPCA_ERR_MISSING, "Unexpected missing genotype in SnpBlockData (prepare.rs:1910)");
  return GPCA_OK;
}

// ---- sketch passes ---------------------------------------------------------------------------------
// Device times of the passes are collected only on request (gpca_set_sketch_timing): a long-lived host that never polls
// gpca_sketch_stats must not accumulate events.
static int timing_begin(gpca_ctx* c, cudaEvent_t* e0, cudaEvent_t* e1) {
  *e0 = *e1 = nullptr;
  if (!c->sk_timing) return GPCA_OK;
  if (c->pending_events.size() > 8192) {     // never polled: fold what has finished into the totals
    GPCA_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    double a = 0, b = 0;
    uint64_t n = 0;
    GPCA_TRY(gpca_sketch_stats(c, &a, &b, &n, 0));
  }
  GPCA_CUDA_TRY(c, cudaEventCreate(e0));
  GPCA_CUDA_TRY(c, cudaEventCreate(e1));
  GPCA_CUDA_TRY(c, cudaEventRecord(*e0, c->stream));
  return GPCA_OK;
}
static int timing_end(gpca_ctx* c, cudaEvent_t e0, cudaEvent_t e1) {
  if (!e0) return GPCA_OK;
  GPCA_CUDA_TRY(c, cudaEventRecord(e1, c->stream));
  c->pending_events.push_back({e0, e1});
  return GPCA_OK;
}

int timed_sketch(gpca_ctx* c, const SketchProblem& p) {
  cudaEvent_t e0, e1;
  GPCA_TRY(timing_begin(c, &e0, &e1));
  const int rc = launch_sketch(c, p);
  GPCA_TRY(timing_end(c, e0, e1));
  c->sk_bytes += (double)p.G.rows * (double)((p.G.cols + 3) / 4);
  c->sk_passes += 1;
  return rc;
}

int timed_sketch_batch(gpca_ctx* c, const SketchBatch& sb) {
  cudaEvent_t e0, e1;
  GPCA_TRY(timing_begin(c, &e0, &e1));
  const int rc = launch_sketch_i8_batch(c, sb);
  GPCA_TRY(timing_end(c, e0, e1));
  c->sk_bytes += sb.bytes;
  c->sk_passes += 1;
  return rc;
}

// A pass over the SNP-major matrix, segment by segment: the resident rows as they are, the rest re-created window by
// window from the sample-major matrix (gpca_ctx::gs_res_rows).  fn(view, first logical row) runs on each segment.
int for_each_gs_segment(gpca_ctx* c, const std::function<int(const PackedMat&, uint64_t)>& fn) {
  const uint64_t D = c->D;
  const uint64_t res = c->gs_win_rows ? std::min<uint64_t>(c->gs_res_rows, D) : D;
  if (res) {
    PackedMat v = c->Gs;
    v.rows = res;
    v.avail = c->Gs.pitch;
    GPCA_TRY(fn(v, 0));
  }
  const bool trace = getenv("GPCA_TRACE") != nullptr && res < D;
  cudaEvent_t te[3] = {nullptr, nullptr, nullptr};
  float t_tr = 0.f, t_fn = 0.f;
  if (trace)
    for (auto& e : te) cudaEventCreate(&e);
  for (uint64_t r = res; r < D; r += c->gs_win_rows) {
    const uint64_t nr = std::min<uint64_t>(c->gs_win_rows, D - r);
    if (trace) cudaEventRecord(te[0], c->stream);
    PackedMat sv = c->Gt, dv = c->Gs;
    sv.p = c->Gt.p + r / 4;               // r is a multiple of 512 (res and the window are)
    sv.cols = nr;
    dv.p = c->Gs.p + c->gs_res_rows * c->Gs.pitch;
    dv.rows = nr;
    dv.avail = c->Gs.pitch;
    // the transpose reads Gt columns [r, r + nr) -- the last 512-column tile may run past nr inside Gt's row: those
    // fields are other SNPs' or zero pads, and land in window rows past nr that the pass does not read
    GPCA_TRY(launch_transpose(c, sv, dv));
    if (trace) cudaEventRecord(te[1], c->stream);
    GPCA_TRY(fn(dv, r));
    if (trace) {
      cudaEventRecord(te[2], c->stream);
      cudaEventSynchronize(te[2]);
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, te[0], te[1]);
      cudaEventElapsedTime(&b, te[1], te[2]);
      t_tr += a;
      t_fn += b;
    }
  }
  if (trace) {
    fprintf(stderr, "[for_each_gs_segment] %llu of %llu rows re-created from the sample-major matrix: transposes %.2f ms, "
            "passes on the windows %.2f ms\n", (unsigned long long)(D - res), (unsigned long long)D, t_tr, t_fn);
    for (auto& e : te) cudaEventDestroy(e);
  }
  return GPCA_OK;
}

int sketch_snp_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld_in, uint32_t ld_out,
                    bool emit_stats, bool out_pad) {
  const bool whole = c->gs_win_rows == 0 || c->gs_res_rows >= c->D;
  return for_each_gs_segment(c, [&](const PackedMat& g, uint64_t row0) -> int {
    SketchProblem p;
    p.out_pad = out_pad;
    p.emit_stats = emit_stats && whole;   // (the statistic is the max-abs over the l logical columns: independent of the stride)
    p.G = g;
    p.Bin = dev_in;
    p.l = l;
    p.ld = ld_in;
    p.f = nullptr;
    p.e = nullptr;
    p.a = c->d_inv_sd.p + row0;
    p.b = c->d_mu_inv_sd.p + row0;
    p.out = dev_out + row0 * ld_out;
    p.ldo = ld_out;
    return timed_sketch(c, p);
  });
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
  if (c->sharded()) {
    if (ld_out != l) return fail(c, GPCA_ERR_INVALID, "sharded sample-side sketch needs ld == l");
    GPCA_TRY(driver_allreduce(c, dev_out, c->N * (uint64_t)l, 0));
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
  if (c->sharded()) {
    if (ld_out != l) return fail(c, GPCA_ERR_INVALID, "sharded sample-side sketch needs ld == l");
    GPCA_TRY(driver_allreduce(c, dev_out, c->N * (uint64_t)l, 0));
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

extern "C" int gpca_dense_product(gpca_ctx* c, const float* dev_c, uint64_t n, uint64_t r, uint32_t ldc, int cols_mode,
                                  const float* dev_in, uint32_t l, uint32_t ld, const float* dev_f, const float* dev_e,
                                  const float* dev_a, const float* dev_b, float* dev_out, uint32_t ldo) {
  CHECK_CTX(c);
  GPCA_CUDA_TRY(c, cudaSetDevice(c->device));
  if (!dev_c || !dev_in || !dev_out || l == 0 || l > 64 || ld < l || ldo < l || ldc < r)
    return fail(c, GPCA_ERR_INVALID, "gpca_dense_product: bad argument");
  DenseProduct dp{dev_c, n, r, ldc, cols_mode != 0, dev_in, l, ld, dev_f, dev_e, dev_a, dev_b, dev_out, ldo};
  return launch_dense_product(c, dp);
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
  const bool trace_launches = getenv("GPCA_TRACE_SKETCH") != nullptr;
  for (size_t i = 0; i < c->pending_kernel_events.size(); ++i) {
    auto& pr = c->pending_kernel_events[i];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) c->sk_kernel_ms += ms;
    if (trace_launches && i < c->pending_kernel_notes.size()) {
      const auto& nt = c->pending_kernel_notes[i];
      const double gb = (double)nt.rows * (double)((nt.K + 3) / 4) * 1e-9;
      fprintf(stderr, "[sketch kernel] rows %llu K %llu ksplit %u %s items %u: %.3f ms  %.1f GB/s\n",
              (unsigned long long)nt.rows, (unsigned long long)nt.K, nt.ksplit & 0xffffu,
              (nt.ksplit & 0x10000u) ? "RT4" : "RT2", nt.items, ms, ms > 0 ? gb / (ms * 1e-3) : 0.0);
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  c->pending_kernel_events.clear();
  c->pending_kernel_notes.clear();
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
