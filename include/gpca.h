/*
 * gpca.h -- C ABI of the B200-native hot path of genomic_pca.
 *
 * This is the drop-in boundary: exactly the calls the reference's Rust host would bind
 * through an FFI crate in place of (a) efficient_pca::PCA::{rfit,transform} and
 * (b) the PcaReadyGenotypeAccessor / EigenSNPCoreAlgorithm::compute_pca pair.
 * Plain pointers and sizes only; no C++/torch types.  INTEGRATION.md shows the Rust
 * `extern "C"` block and the two call-site patches.
 *
 * Conventions:
 *   - every function returns 0 on success, <0 on error; gpca_last_error(ctx) gives text;
 *   - the caller allocates and frees every output buffer; the library owns device memory,
 *     streams and events inside the opaque context;
 *   - one context per GPU per process; calls on one context are serialized by the caller;
 *   - there is NO CPU fallback: every entry point that computes needs an sm_100 device.
 *   - "host" pointers are ordinary host memory (pinned or pageable); "dev" pointers are
 *     CUDA device pointers on the context's device.
 *
 * Layouts:
 *   - PLINK payload: SNP-major, row = ceil(N/4) bytes, sample i in bits 2*(i%4) of byte i/4,
 *     codes 00 -> dosage 2 (hom A1), 01 -> missing, 10 -> 1, 11 -> 0  (bed-reader count_a1,
 *     reference src/prepare.rs:622-629).
 *   - all dense matrices are row-major.
 */
#ifndef GPCA_H_
#define GPCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpca_ctx gpca_ctx;
struct gpca_eigensnp_cfg_s;

enum {
  GPCA_OK = 0,
  GPCA_ERR_INVALID = -1,   /* bad argument / wrong call order */
  GPCA_ERR_CUDA = -2,      /* CUDA runtime error (text in gpca_last_error) */
  GPCA_ERR_NO_DEVICE = -3, /* no sm_100 device: there is no CPU fallback */
  GPCA_ERR_MISSING = -4,   /* missing genotype where the reference would error (prepare.rs:1906-1912) */
  GPCA_ERR_OOM = -5
};

/* ---- lifetime ------------------------------------------------------------------------ */
int gpca_init(gpca_ctx** ctx, int device);
void gpca_destroy(gpca_ctx* ctx);
const char* gpca_last_error(const gpca_ctx* ctx);
const char* gpca_version(void);
/* kernels launched by this context since creation / last reset (bench.py "gpu_launches") */
uint64_t gpca_launch_count(const gpca_ctx* ctx);
void gpca_reset_launch_count(gpca_ctx* ctx);
/* 0 = SIMT fp32 path, 1 = tcgen05 f16 path, 2 = tcgen05 i8 path (exact integer accumulation, l <= 32);
 * engines 1/2 fall back to the next lower one for shapes they do not take */
int gpca_set_sketch_engine(gpca_ctx* ctx, int engine);
/* the engine the last (non-batched) sketch pass actually ran on: 0 / 1 / 2, -1 before the first pass */
int gpca_last_sketch_engine(const gpca_ctx* ctx);
/* EigenSNP local bases / condensed features: 1 (default) = every LD block in one launch per stage (integer engine,
 * no missing calls; other cases use the per-block path automatically), 0 = always one block at a time.
 * Replaces the per-block loop inside EigenSNPCoreAlgorithm::compute_pca (call site src/main.rs:365). */
int gpca_set_batch_blocks(gpca_ctx* ctx, int on);
/* device time (ms) and algorithmic packed bytes of the sketch passes since the last reset */
int gpca_sketch_stats(gpca_ctx* ctx, double* ms_total, double* packed_bytes_total, uint64_t* n_passes, int reset);
/* device time (ms) of the main sketch kernel launches alone (no operand prep / split-K reduce), as of the last
 * gpca_sketch_stats call */
double gpca_sketch_kernel_ms(gpca_ctx* ctx);

/* Host threads this context may use for its host-side stages (QC ladder, compaction, widening of results): 0 = every
 * CPU the process may run on (default; GPCA_HOST_THREADS overrides).  With one process per GPU the host sets
 * cores / ranks so that the ranks do not oversubscribe the machine.  Stands in for `--threads` of the reference CLI
 * (src/main.rs:103-106, the global rayon pool). */
int gpca_set_host_threads(gpca_ctx* ctx, uint32_t n_threads);
/* Binds the calling thread -- and the threads created from it afterwards (the context's host pool) -- to the CPUs that
 * sysfs lists next to the context's GPU (/sys/bus/pci/devices/<bus id>/local_cpulist), so that pinned payload buffers
 * first touched afterwards (gpca_host_alloc) are NUMA-local to the GPU.  Call it once, right after gpca_init, from the
 * thread that drives the context.  Returns the number of CPUs in the set (0: not available, nothing changed). */
int gpca_bind_host_to_device(gpca_ctx* ctx);
/* 1 = record CUDA events around every sketch pass so that gpca_sketch_stats / gpca_sketch_kernel_ms report device
 * times (benchmarks); 0 (default) = no events are created on the production path. */
int gpca_set_sketch_timing(gpca_ctx* ctx, int on);
/* Device bytes the next ingest must leave free for the drivers' working buffers (default 6 GiB; gpca_eigensnp needs
 * about gpca_eigensnp_workspace_bytes).  When both orientations of the packed matrix plus this reserve do not fit, the
 * ingest keeps the sample-major copy whole and only the first rows of the SNP-major copy resident
 * (gpca_resident_snp_rows); the passes re-create the rest window by window. */
int gpca_set_memory_reserve(gpca_ctx* ctx, uint64_t bytes);
uint64_t gpca_resident_snp_rows(const gpca_ctx* ctx);
/* Working memory gpca_eigensnp needs on the device for n_samples x n_pca_snps in n_blocks LD blocks (an upper
 * estimate; pass it to gpca_set_memory_reserve before the ingest). */
uint64_t gpca_eigensnp_workspace_bytes(uint64_t n_samples, uint64_t n_pca_snps, uint64_t n_blocks,
                                       const struct gpca_eigensnp_cfg_s* cfg);

/* Pinned host memory for large payloads (a host that reads the .bed itself and hands it to gpca_ingest_bed): anonymous
 * memory on transparent huge pages, first-touched on the context's host threads and registered with CUDA -- several
 * times faster to obtain than cudaMallocHost for tens of GB.  Stands in for the reference's Mmap of the .bed
 * (bed-reader, src/prepare.rs:491). */
void* gpca_host_alloc(gpca_ctx* ctx, uint64_t bytes);
void gpca_host_free(gpca_ctx* ctx, void* p, uint64_t bytes);

/* ---- exchange between shards (multi-GPU) ------------------------------------------------ */
/* The library's own communicator: NCCL over NVLink / NVSwitch, bound at run time (libnccl.so.2).  One context = one
 * rank; the contexts may live in threads of one process or in one process each.  One rank calls gpca_comm_unique_id,
 * the host passes the id bytes to the others (MPI, a file, torch.distributed ...), every rank calls gpca_comm_init;
 * from then on the library issues its collectives (sum of the N x l sketch after a sample-side pass, l x l Grams) on
 * its own stream.  Nothing in the reference corresponds to this (single process). */
#define GPCA_COMM_ID_BYTES 128
int gpca_comm_unique_id(uint8_t* id_out /* GPCA_COMM_ID_BYTES */);
int gpca_comm_init(gpca_ctx* ctx, const uint8_t* id /* GPCA_COMM_ID_BYTES */, int rank, int world);
int gpca_comm_finalize(gpca_ctx* ctx);
int gpca_comm_world(const gpca_ctx* ctx);
/* collectives issued by this context since creation (either transport) */
uint64_t gpca_collective_count(const gpca_ctx* ctx);

/* Cross-shard sum hook (multi-GPU, SNP-sharded): called on the context's stream order with a
 * DEVICE buffer that must be replaced by its sum over all shards (fp32 or fp64).
 * dtype: 0 = f32, 1 = f64.  Replaces nothing in the reference (single process); it is where
 * the host binds ncclAllReduce.  NULL (default) = single shard. */
typedef int (*gpca_allreduce_fn)(void* dev_buf, uint64_t count, int dtype, void* cuda_stream, void* user);
int gpca_set_allreduce(gpca_ctx* ctx, gpca_allreduce_fn fn, void* user);
/* global row offset / total of this shard's variants (for shard-independent random streams) */
int gpca_set_shard(gpca_ctx* ctx, uint64_t variant_offset, uint64_t variants_total);

/* ---- ingest -------------------------------------------------------------------------- */
/* Replaces IoService + bed-reader reads (src/prepare.rs:169-920, 606-629): the packed payload
 * (after the 3-byte magic) is copied to the device once.  keep_samples (original FAM indices,
 * increasing; src/prepare.rs:1058-1096) may be NULL = all samples. */
int gpca_load_bed(gpca_ctx* ctx, const uint8_t* host_payload, uint64_t n_samples_in_file, uint64_t n_snps,
                  const int64_t* keep_samples, uint64_t n_keep);
/* Same, payload already on the device (row pitch = ceil(n_samples/4)); no sample subsetting. */
int gpca_load_bed_device(gpca_ctx* ctx, const uint8_t* dev_payload, uint64_t n_samples, uint64_t n_snps);
/* Replaces vcf::matrix_ops::build_matrix (src/vcf.rs:317-345): variant-major u8 dosages
 * (0,1,2; any other value = missing), D rows of N bytes, as aggregated at src/vcf.rs:293-315. */
int gpca_load_u8_variant_major(gpca_ctx* ctx, const uint8_t* host_dosage, uint64_t n_samples, uint64_t n_variants);

uint64_t gpca_num_samples(const gpca_ctx* ctx);
uint64_t gpca_num_snps(const gpca_ctx* ctx);      /* loaded rows (before QC) */
uint64_t gpca_num_pca_snps(const gpca_ctx* ctx);  /* after gpca_set_pca_snps */

/* ---- statistics / QC ----------------------------------------------------------------- */
/* Pass 1 of perform_snp_qc_and_calc_std_params (src/prepare.rs:1232-1279): per loaded SNP
 * n_valid, n(dosage 0), n(dosage 1), n(dosage 2) over the kept samples.  Bit-exact integers. */
int gpca_snp_counts(gpca_ctx* ctx, uint32_t* n_valid, uint32_t* n0, uint32_t* n1, uint32_t* n2);

typedef struct {
  double min_call_rate; /* src/main.rs:545  default 0.98 */
  double min_maf;       /* src/main.rs:548  default 0.01 */
  double max_hwe_p;     /* src/main.rs:551  default 1e-6 ; >= 1.0 disables */
} gpca_qc_cfg;
/* The QC ladder + mean/sigma (src/prepare.rs:1281-1375) in f64 on the host from the integer
 * counts.  keep[M] (0/1), mean[M], sd[M] (f32, 0 where dropped), fail_code[M] may be NULL.
 * fail codes: 0 kept, 1 call-rate, 2 no valid, 3 maf, 4 monomorphic, 5 hwe, 6 variance. */
int gpca_snp_qc(gpca_ctx* ctx, const gpca_qc_cfg* cfg, uint8_t* keep, float* mean, float* sd, uint8_t* fail_code);
/* VCF-mode filter (src/vcf.rs:244-266): keep iff min(p,1-p) >= maf, p = sum/(2N); variants with
 * any missing call are dropped (src/vcf.rs:227-242).  mean/sd: column mean, ddof=1 sd
 * (sd <= 1e-9 -> 1) as rfit standardises. */
int gpca_vcf_maf_filter(gpca_ctx* ctx, double maf_threshold, uint8_t* keep, float* mean, float* sd);
/* HWE chi-square p-value alone (src/prepare.rs:1641-1745); pure host arithmetic, exported so
 * the parity tests can pin it. */
double gpca_hwe_chi_squared_p_value(uint64_t hom1, uint64_t het, uint64_t hom2);

/* Select the PCA SNP set (PcaSnpId i <-> loaded row snp_idx[i], strictly increasing) with its
 * standardisation parameters; builds the resident device copies used by every sketch pass.
 * Mirrors MicroarrayGenotypeAccessor::new (src/prepare.rs:1783-1822). */
int gpca_set_pca_snps(gpca_ctx* ctx, const uint64_t* snp_idx, uint64_t n_pca_snps, const float* mean, const float* sd);
/* After gpca_ingest_bed (which keeps no staging copy of the payload) both calls can only NARROW the resident set: the
 * kept rows move up inside the resident matrices.  This is the EigenSNP workflow's step that drops QC'd SNPs outside
 * every LD block (src/prepare.rs:1424-1563); hosts that know the selection up front pass it to the ingest instead
 * (gpca_set_ingest_mask). */
/* Same, selecting every loaded SNP with keep[j] != 0 and taking mean/sd from the full-length arrays that
 * gpca_snp_qc / gpca_vcf_maf_filter filled (saves the host a gather over millions of SNPs). */
int gpca_set_pca_snps_mask(gpca_ctx* ctx, const uint8_t* keep, const float* mean_all, const float* sd_all,
                           uint64_t* n_pca_out);
/* One streaming pass that is equivalent to gpca_load_bed + gpca_snp_qc (cfg != NULL) or gpca_vcf_maf_filter
 * (cfg == NULL, threshold = vcf_maf_threshold) + gpca_set_pca_snps_mask: while later chunks of the payload are still
 * crossing PCIe, earlier chunks are counted on the device, filtered on host threads and recoded into the resident
 * matrix.  keep / mean / sd / fail_code (n_snps each) may be NULL.  Stands in for the whole data-preparation stage
 * (src/prepare.rs:995-1098 for .bed input; src/vcf.rs:227-266 + src/main.rs:176-212 for the VCF flow). */
int gpca_ingest_bed(gpca_ctx* ctx, const uint8_t* host_payload, uint64_t n_in_samples, uint64_t n_snps,
                    const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg, double vcf_maf_threshold,
                    uint8_t* keep, float* mean, float* sd, uint8_t* fail_code, uint64_t* n_pca_out);
/* Optional pre-selection for the next gpca_ingest_bed / gpca_ingest_bed_file: loaded rows with mask[j] == 0 are dropped
 * whatever the QC says (fail code 7) -- e.g. SNPs that lie in no LD block (src/prepare.rs:1447-1463 depends only on
 * chromosome / position, so the host can evaluate it before the genotypes are read).  NULL clears it. */
int gpca_set_ingest_mask(gpca_ctx* ctx, const uint8_t* mask, uint64_t n_snps);
/* The same pass straight from a PLINK .bed file (magic and size are checked against n_in_samples x n_snps): chunks are
 * read into pinned buffers and copied on, so the file is never held in host memory as a whole and the read overlaps
 * with the transfer and the device work.  Replaces the IoService reader pool for this stage (src/prepare.rs:169-920,
 * :923-993). */
int gpca_ingest_bed_file(gpca_ctx* ctx, const char* bed_path, uint64_t n_in_samples, uint64_t n_snps,
                         const int64_t* keep_samples, uint64_t n_keep, const gpca_qc_cfg* cfg, double vcf_maf_threshold,
                         uint8_t* keep, float* mean, float* sd, uint8_t* fail_code, uint64_t* n_pca_out);

/* Rows [first_row, first_row + n_rows) of the file only -- the shard of one GPU of a multi-GPU run; keep / mean / sd /
 * fail_code (n_rows each) and the ingest mask are indexed by the row inside the range. */
int gpca_ingest_bed_file_rows(gpca_ctx* ctx, const char* bed_path, uint64_t n_in_samples, uint64_t n_snps_in_file,
                              uint64_t first_row, uint64_t n_rows, const int64_t* keep_samples, uint64_t n_keep,
                              const gpca_qc_cfg* cfg, double vcf_maf_threshold, uint8_t* keep, float* mean, float* sd,
                              uint8_t* fail_code, uint64_t* n_pca_out);

/* ---- the accessor the GPU path makes unnecessary, kept for parity ---------------------- */
/* get_standardized_snp_sample_block (src/prepare.rs:1839-2022): out[n_ids x n_samp] row-major
 * f32, z = fma(x, 1/sd, -mean/sd); sd < 1e-9 -> 0; any missing call -> GPCA_ERR_MISSING. */
int gpca_get_standardized_block(gpca_ctx* ctx, const uint64_t* pca_snp_ids, uint64_t n_ids,
                                const uint64_t* qc_sample_ids, uint64_t n_samp, float* host_out);

/* ---- sketch passes (the hot path), device-pointer level -------------------------------- */
/* S = standardized [D x N] (never materialised).  Missing calls contribute 0.
 *   snp side   : dev_out[D x l]  = S   * dev_in[N x l]      (reduction over samples)
 *   sample side: dev_out[N x l]  = S^T * dev_in[D x l]      (reduction over SNPs; summed over
 *                                                            shards through the exchange)
 * l <= 64.  ld = row stride in floats of both dense operands.
 * On a sharded context (gpca_set_shard + gpca_comm_init / gpca_set_allreduce) the sample-side call is a
 * COLLECTIVE: it ends with the allreduce of the N x l result and needs ld == l; every shard must issue it,
 * in the same order as its other collectives.  The snp-side call is local to the shard. */
int gpca_sketch_snp_side(gpca_ctx* ctx, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld);
int gpca_sketch_sample_side(gpca_ctx* ctx, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld);
/* Product with a dense fp32 matrix C [n x r] (row stride ldc floats) on the device -- the condensed-feature matrix of
 * EigenSNP's global randomized SVD (configured at src/main.rs:317-318), exported so that the kernel can be tested alone:
 *   cols_mode = 0: dev_out[n x l] = a o (C   (f o dev_in[r x l])) - b (x) (e^T dev_in)     (rows = samples)
 *   cols_mode = 1: dev_out[r x l] = a o (C^T (f o dev_in[n x l])) - b (x) (e^T dev_in)     (rows = columns of C)
 * f / e index the reduction axis, a / b the output rows; any of them may be NULL (= 1).  Split-bf16 tensor-core
 * arithmetic (about 2^-16 relative) for l <= 32 and ldc % 4 == 0, plain fp32/f64 otherwise. */
int gpca_dense_product(gpca_ctx* ctx, const float* dev_c, uint64_t n, uint64_t r, uint32_t ldc, int cols_mode,
                       const float* dev_in, uint32_t l, uint32_t ld, const float* dev_f, const float* dev_e,
                       const float* dev_a, const float* dev_b, float* dev_out, uint32_t ldo);
int gpca_synchronize(gpca_ctx* ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on -- for CUDA-event timing */
void* gpca_get_stream(gpca_ctx* ctx);

/* ---- drivers --------------------------------------------------------------------------- */
/* Replaces pca_runner::run_genomic_pca = PCA::rfit + PCA::transform (src/main.rs:598-679).
 * scores[N x k_out] f64 row-major (the reference's Array2<f64>), explained variance
 * eigenvalues[k_out] = s^2/(N-1), loadings[D x k_out] f32 (rotation; may be NULL).
 * oversample: src/main.rs:636 (10).  has_seed=0 mirrors `--rfit-seed` absent (entropy seed; with the library's
 * communicator shard 0's seed is broadcast, with the host hook an explicit seed is required).
 * scores / eigenvalues may be NULL: the shards of a multi-GPU run all hold the same scores, and a host that reads
 * them from one shard passes NULL on the others (no download, no widening to f64 there). */
int gpca_rfit(gpca_ctx* ctx, uint32_t k, uint32_t oversample, uint32_t power_iters, uint64_t seed, int has_seed,
              double* scores, double* eigenvalues, float* loadings, uint32_t* k_out);

typedef struct gpca_eigensnp_cfg_s {     /* EigenSNPCoreAlgorithmConfig, src/main.rs:311-327 */
  uint32_t target_num_global_pcs;        /* --eigensnp-k-global            10    */
  uint32_t components_per_ld_block;      /* --eigensnp-components-per-block 7    */
  double subset_factor;                  /* --eigensnp-subset-factor       0.075 */
  uint64_t min_subset_size;              /* --eigensnp-min-subset-size     10000 */
  uint64_t max_subset_size;              /* --eigensnp-max-subset-size     40000 */
  uint32_t global_oversampling;          /* --eigensnp-global-oversampling 10    */
  uint32_t global_power_iters;           /* --eigensnp-global-power-iter   2     */
  uint32_t local_oversampling;           /* --eigensnp-local-oversampling  10    */
  uint32_t local_power_iters;            /* --eigensnp-local-power-iter    2     */
  uint64_t random_seed;                  /* --eigensnp-seed                2025  */
  uint32_t snp_processing_strip_size;    /* --eigensnp-snp-strip-size      2000 (accepted, unused: no strips on GPU) */
  uint32_t refine_pass_count;            /* --eigensnp-refine-passes       1     */
  uint32_t collect_diagnostics;          /* --eigensnp-collect-diagnostics: keep a JSON record of the run (below) */
} gpca_eigensnp_cfg;
void gpca_eigensnp_default_cfg(gpca_eigensnp_cfg* cfg);

/* Replaces EigenSNPCoreAlgorithm::compute_pca(&accessor, &ld_block_specifications)
 * (src/main.rs:365).  Blocks in tag-sorted order; block b owns PcaSnpIds
 * block_snp_ids[block_offsets[b] .. block_offsets[b+1]) (sorted within a block),
 * i.e. Vec<LdBlockSpecification> flattened (src/prepare.rs:1526-1549).
 * scores[N x k_out] f32, eigenvalues[k_out] f64, loadings[D x k_out] f32
 * (final_sample_principal_component_scores / _eigenvalues / final_snp_principal_component_loadings).
 * scores / eigenvalues / loadings may be NULL (see gpca_rfit). */
int gpca_eigensnp(gpca_ctx* ctx, const gpca_eigensnp_cfg* cfg, const uint64_t* block_offsets, uint64_t n_blocks,
                  const uint64_t* block_snp_ids, float* scores, double* eigenvalues, float* loadings, uint32_t* k_out);

/* Diagnostics of the last gpca_eigensnp call made with cfg->collect_diagnostics != 0, as a JSON document (shape of the
 * run, layout decisions, per-stage device times, launches, collectives, eigenvalues); "" if none.  Stands in for the
 * `<prefix>.eigensnp_diagnostics.json` the reference writes behind its eigensnp-diagnostics feature
 * (src/main.rs:411-430) -- the schema of the external crate's FullPcaRunDetailedDiagnostics is not known here, so this
 * is the library's own.  The pointer is valid until the next gpca_eigensnp / gpca_destroy on the context. */
const char* gpca_eigensnp_diagnostics(const gpca_ctx* ctx);

/* ---- host-side helpers of the path (no GPU needed) -------------------------------------- */
/* map_snps_to_ld_blocks (src/prepare.rs:1424-1563) over parsed blocks.  Inputs per QC'd SNP
 * in increasing original index; chromosomes as normalised strings.  Outputs: pca_pos[n_qc]
 * = PcaSnpId or -1; block_of[n_qc] = index into the tag-sorted block list or -1;
 * returns number of PCA SNPs through n_pca and number of non-empty blocks through n_blocks_out;
 * sorted_block_order[n_blocks] maps sorted position -> input block index. */
int gpca_map_snps_to_ld_blocks(const char* const* snp_chrom, const int32_t* snp_bp, uint64_t n_qc,
                               const char* const* blk_chrom, const int32_t* blk_start, const int32_t* blk_end,
                               uint64_t n_blocks, int64_t* pca_pos, int64_t* block_of, uint64_t* n_pca,
                               uint64_t* n_blocks_out, uint64_t* sorted_block_order);

/* Benchmark input (not on the reference's path): synthetic structured genotypes written on the
 * device straight in .bed layout (n_snps rows of ceil(n_samples/4) bytes).  Counter-based (Philox keyed by seed, global
 * SNP index = snp_offset + row, sample), so any shard of SNPs regenerates identically at any GPU count.
 * fst_grade > 0 gives population k the drift fst * (1 + fst_grade * (0.5 - k / (n_pops - 1))): distinct structural
 * eigenvalues, so that a top-k subspace is well defined for any k (0 = one F_ST for all: a degenerate cluster). */
int gpca_synth_bed_device(gpca_ctx* ctx, uint8_t* dev_out, uint64_t n_samples, uint64_t n_snps, uint64_t snp_offset,
                          uint64_t seed, uint32_t n_pops, double fst, double missing_rate, double fst_grade);
/* Measurement hook: mean device time (ms) of the ingest's allele-count kernel (K-a: 2-bit unpack + counts, the integer
 * half of src/prepare.rs:1232-1279) over `reps` launches on a device-resident .bed payload of n_snps rows of
 * ceil(n_samples/4) bytes (the payload must be readable 16 bytes past its end). */
int gpca_count_kernel_ms(gpca_ctx* ctx, const uint8_t* dev_payload, uint64_t n_samples, uint64_t n_snps, uint32_t reps,
                         double* ms_out);
/* the same rows into host memory (generated on the device chunk by chunk): stands in for a .bed file read by the host */
int gpca_synth_bed_host(gpca_ctx* ctx, uint8_t* host_out, uint64_t n_samples, uint64_t n_snps, uint64_t snp_offset,
                        uint64_t seed, uint32_t n_pops, double fst, double missing_rate, double fst_grade);

#ifdef __cplusplus
}
#endif
#endif /* GPCA_H_ */
