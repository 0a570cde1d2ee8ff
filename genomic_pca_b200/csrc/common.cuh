// common.cuh -- context, error plumbing and small device helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gpca.h"
#include "parallel_for.h"

#define GPCA_CUDA_TRY(ctx, expr)                                                         \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
      return (_e == cudaErrorMemoryAllocation) ? GPCA_ERR_OOM : GPCA_ERR_CUDA;           \
    }                                                                                    \
  } while (0)

#define GPCA_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc != GPCA_OK) return _rc; \
  } while (0)

// Owning device buffer (raw cudaMalloc; freed in dtor).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

// Call-scoped device temporaries that survive between calls: a driver asks for its buffers in the same order on
// every call, so the k-th request reuses the k-th buffer (grown when needed) and no cudaMalloc / cudaFree remains on
// the path of a repeated call.
struct ScratchPool {
  std::vector<DevBuf<uint8_t>*> bufs;
  size_t next = 0;
  ~ScratchPool() { release(); }
  void reset() { next = 0; }
  size_t bytes() const {
    size_t t = 0;
    for (auto* b : bufs) t += b->n;
    return t;
  }
  void release() {
    for (auto* b : bufs) delete b;
    bufs.clear();
    next = 0;
  }
  cudaError_t get(size_t bytes, void** out) {
    if (next >= bufs.size()) bufs.push_back(new DevBuf<uint8_t>());
    DevBuf<uint8_t>* b = bufs[next++];
    const cudaError_t e = b->alloc(bytes ? bytes : 1);
    *out = b->p;
    return e;
  }
};
template <typename T>
struct PoolBuf {       // DevBuf-shaped view of a pool buffer
  ScratchPool* pool;
  T* p = nullptr;
  size_t n = 0;
  explicit PoolBuf(ScratchPool* pl) : pool(pl) {}
  cudaError_t alloc(size_t count) {
    void* q = nullptr;
    const cudaError_t e = pool->get(count * sizeof(T), &q);
    p = static_cast<T*>(q);
    n = (e == cudaSuccess) ? count : 0;
    return e;
  }
};

static inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// Resident packed genotype matrix, dosage-coded 2-bit fields (0,1,2 = A1 dosage, 3 = missing),
// field k of a row in bits 2*(k%4) of byte k/4, pad fields = 0.  pitch is a multiple of 128 B.
struct PackedMat {
  uint8_t* p = nullptr;
  size_t pitch = 0;      // bytes per row
  uint64_t rows = 0;     // logical rows
  uint64_t cols = 0;     // logical 2-bit fields per row
  size_t avail = 0;      // bytes readable from p to the end of the physical row (0 = pitch); views set it
};

struct gpca_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // H2D transfers of the streaming ingest (so that they do not queue behind kernels)
  std::string err;
  uint64_t launches = 0;
  int engine = 2;   // 0 SIMT fp32, 1 tcgen05 f16, 2 tcgen05 i8 (default; l > 32 falls back to 1)
  int last_engine = -1;   // the engine the last regular sketch pass actually ran on
  int batch_blocks = 1;   // EigenSNP: all LD blocks per launch (needs engine 2); 0 = one block at a time

  // exchange between shards: the library's own NCCL communicator (gpca_comm_init, comm.cu) or a host-provided hook
  gpca_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  void* nccl_comm = nullptr;           // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  uint64_t collectives = 0;            // collectives issued since creation (bench / tests)
  uint64_t shard_vote_D = 0;           // gpca_rfit: the shard size the cached side decision was made for
  int shard_vote_snp_side = -1;
  uint64_t shard_offset = 0, shard_total = 0;
  bool sharded() const { return allreduce != nullptr || nccl_comm != nullptr; }

  // loaded (pre-QC) data, PLINK-coded, kept samples only, pad fields = 01
  uint64_t N = 0, M = 0;
  DevBuf<uint8_t> raw;
  size_t raw_pitch = 0;
  bool have_counts = false;
  std::vector<uint32_t> h_counts;  // [M][4] = n_valid, n0, n1, n2
  bool vcf_mode = false;

  // PCA SNP set + resident copies
  uint64_t D = 0;
  std::vector<uint64_t> pca_idx;
  std::vector<float> h_mean, h_sd, h_inv, h_muinv;
  float inv_sd_max = 0.f;      // max 1/sd over the PCA SNPs (bounds |f o Omega| for the generated test matrix)
  DevBuf<uint64_t> d_idx;
  DevBuf<uint4> d_cnt;
  // count records of gpca_snp_counts / the streaming ingest: pinned AND mapped, so that the ingest's count kernel
  // writes them straight into host memory (h_cnt_dev = the device-side address of the same pages)
  uint4* h_cnt = nullptr;
  uint4* h_cnt_dev = nullptr;
  uint64_t h_cnt_cap = 0;
  DevBuf<uint8_t> es_store, et_store, ets_store, ess_store;   // EigenSNP slot-ordered / subset copies (kept across calls)
  ScratchPool es_pool;                                        // EigenSNP call-scoped temporaries
  std::vector<int64_t> es_subset;                             // the N_s-sample subset of the last EigenSNP call
  uint64_t es_subset_n = 0, es_subset_seed = 0;
  // every change of the resident matrices bumps data_version; the subset copies of EigenSNP remember which one they were made from
  uint64_t data_version = 1, es_sub_copies_version = 0, es_sub_copies_ns = 0, es_sub_copies_seed = 0;
  DevBuf<float> es_cn;                                        // EigenSNP condensed features
  std::string es_diag_json;                                   // diagnostics of the last gpca_eigensnp call that collected them
  // streaming ingest (gpca_ingest_bed): a ring of device staging buffers for the payload chunks, the sample-gathered
  // copy of a chunk when a keep-list is given, pinned+mapped staging of the per-chunk compacted vectors, pinned read
  // buffers of gpca_ingest_bed_file
  static constexpr int INGEST_STAGES = 4;
  DevBuf<uint8_t> ingest_stage[INGEST_STAGES], ingest_gather[INGEST_STAGES];
  uint8_t* h_up = nullptr;
  uint8_t* h_up_dev = nullptr;
  size_t h_up_cap = 0;
  uint8_t* h_rd[INGEST_STAGES] = {nullptr, nullptr, nullptr, nullptr};
  size_t h_rd_cap = 0;
  std::vector<uint8_t> ingest_mask;        // optional pre-selection of loaded rows for the next ingest (gpca_set_ingest_mask)
  size_t mem_reserve = 6ull << 30;         // device bytes the ingest leaves free for the drivers' working buffers
  float* h_dl[2] = {nullptr, nullptr};     // pinned landing buffers of download_results (drivers.cu)
  cudaEvent_t ev_dl[2] = {nullptr, nullptr};
  DevBuf<float> d_mean, d_sd;           // [D]
  DevBuf<float> d_inv_sd, d_mu_inv_sd;  // [D]  1/sd (0 if sd<1e-9) and mean/sd
  DevBuf<uint8_t> gs_store, gt_store;
  PackedMat Gs, Gt;                     // [D x N] and [N x D]
  // Residency of the SNP-major copy.  Both orientations resident is the fast layout; when they do not fit (500,000 x
  // 700,000 on one GPU: 2 x 87.5 GB) Gt stays whole and Gs keeps only its first gs_res_rows rows, followed in the same
  // allocation by a window of gs_win_rows rows: the ingest builds later rows through it as a ring, and a pass over
  // the SNP-major matrix re-creates the non-resident rows window by window from Gt (for_each_gs_segment, api.cu).
  uint64_t gs_res_rows = 0, gs_win_rows = 0;
  bool any_missing = false;

  // sketch statistics (bytes and passes always; device times only when sk_timing is on: gpca_set_sketch_timing)
  bool sk_timing = false;
  double sk_ms = 0, sk_bytes = 0;
  uint64_t sk_passes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_kernel_events;   // around the main sketch kernel only
  struct KernelNote { uint64_t rows, K; uint32_t ksplit, items; };        // what each timed launch covered (GPCA_TRACE_SKETCH)
  std::vector<KernelNote> pending_kernel_notes;
  double sk_kernel_ms = 0, sk_kernel_ms_last = 0;

  // scratch
  DevBuf<float> ws_bprep;      // prepared B' operand
  DevBuf<float> ws_partial;    // split-K partials
  DevBuf<float> ws_cvec;       // epilogue vector(s)
  DevBuf<double> ws_f64;       // f64 output staging
  DevBuf<double> ws_gram;      // gram partials
  DevBuf<double> ws_cpart;     // column-sum partials
  DevBuf<double> ws_small;     // l x l matrices: G, evals, evecs, T
  DevBuf<uint8_t> ws_bytes;    // tcgen05 engine: fp16 B' image
  DevBuf<double> ws_batch;     // batched dense helpers: G / T / evecs / evals / flags per problem
  DevBuf<double> ws_stats;     // operand statistics produced as by-products (see SketchProblem::emit_stats)
  const float* stats_for = nullptr;   // the matrix they describe (nullptr = none)
  uint32_t stats_l = 0;
  int stats_nparts = 0;
  bool stats_pending = false;         // a producer left statistics that no consumer has taken (max-abs word not reset)
  DevBuf<float> ws_bstat;      // batched passes: per-block column sums, scales, amax words
  bool tc_amax_zeroed = false;
  DevBuf<float> drv_a, drv_b, drv_c, drv_d, drv_e;   // driver-level dense operands (kept across calls: no per-call cudaMalloc)

  // host threads of this context (parallel_for.h): a persistent pool, created on first use
  unsigned host_threads = 0;   // 0 = every CPU the process may run on
  std::unique_ptr<HostPool> pool;
  HostPool* host_pool() {
    if (!pool) pool.reset(new HostPool(host_threads ? host_threads : host_cpu_count()));
    return pool.get();
  }

  void set_error(const std::string& s) { err = s; }
};
// installs the context's host-thread pool for the parallel_for calls of one entry point
#define GPCA_HOST_POOL(c) HostPoolScope _host_pool_scope((c)->host_pool())

// CUDA-event bracket around the dominant (sketch) kernel alone: begin() before the launch, end() after it.
struct KernelTimer {
  gpca_ctx* c;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  explicit KernelTimer(gpca_ctx* ctx) : c(ctx) {
    if (!c->sk_timing) return;       // opt-in: a long-lived host never polls gpca_sketch_stats
    if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, c->stream);
  }
  void end(uint64_t rows = 0, uint64_t K = 0, uint32_t ksplit = 0, uint32_t items = 0) {
    if (e0 && e1) {
      cudaEventRecord(e1, c->stream);
      c->pending_kernel_events.push_back({e0, e1});
      c->pending_kernel_notes.push_back({rows, K, ksplit, items});
    }
  }
};

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
