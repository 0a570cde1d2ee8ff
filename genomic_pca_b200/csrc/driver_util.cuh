// driver_util.cuh -- helpers shared by the PCA drivers (drivers.cu, eigensnp.cu).
#pragma once
#include <functional>
#include <vector>

#include "kernels.cuh"
#include "sketch_tc.cuh"

struct Small {  // f64 scratch for l x l work, all on device
  double *G, *evals, *evecs, *T;
  int* flag;   // 1 = the Cholesky transform succeeded (eigen path skipped)
};
int get_small(gpca_ctx* c, Small& s);

// Orthonormalise the columns of Y [n x l] in place: two rounds of  G = Y^T Y = V L V^T ; Y <- Y V L^-1/2
// (eigen-based CholeskyQR2 variant; rank-deficient directions are zeroed instead of breaking a Cholesky).
// `sharded`: rows of Y are split across shards -> the l x l Gram is summed through the allreduce hook.
int orthonormalize(gpca_ctx* c, float* y, uint64_t n, uint32_t l, uint32_t ld, bool sharded, const Small& s);

// t [l x k] = evecs[:, :k] * diag(evals[:k]^-1/2)  (inv_sqrt) or evecs[:, :k] (plain)
int launch_rotation_transform(gpca_ctx* c, const double* evals, const double* evecs, uint32_t l, uint32_t k, double* t,
                              bool inv_sqrt);

void fix_signs_host(std::vector<float>& scores, uint64_t n, uint32_t k, std::vector<int>& flip);

int driver_allreduce(gpca_ctx* c, void* buf, uint64_t count, int dtype);   // comm.cu
int driver_broadcast0(gpca_ctx* c, void* buf, uint64_t bytes);              // comm.cu

// Results back to caller-owned (pageable) host memory: `count` floats from the device arrive in pinned landing buffers
// chunk by chunk, and host threads copy (dst_f32) or widen (dst_f64) chunk q while chunk q+1 is on the bus.  Returns
// after the last element has been written (the stream is idle by then).  A pageable cudaMemcpyAsync of the same data
// ran at ~5 GB/s: 16 ms for the 500,000 x 20 f64 scores of one shard of BASELINE config 4.
int download_results(gpca_ctx* c, const float* d_src, uint64_t count, float* dst_f32, double* dst_f64);

// emit_stats: also leave the operand statistics of the output (as the next sample-side pass needs them) in the context;
// use_stats: the operand's statistics are in the context (it was produced by such a pass or by
// launch_gaussian_with_stats and has not been modified since)
// out_pad: the caller owns columns [l, ld_out) of every output row as padding (they may be overwritten with zeros)
int sketch_snp_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld_in, uint32_t ld_out,
                    bool emit_stats = false, bool out_pad = false);
int sketch_sample_side(gpca_ctx* c, const float* dev_in, float* dev_out, uint32_t l, uint32_t ld_in, uint32_t ld_out,
                       bool use_stats = false);
// Y = S^T Omega with Omega = the D x l Gaussian test matrix of (seed, stream): generated inside the operand preparation
// when the integer engine runs (dev_scratch [D x ld_in] is only written by the other paths)
int sketch_sample_side_gaussian(gpca_ctx* c, float* dev_scratch, float* dev_out, uint32_t l, uint32_t ld_in,
                                uint32_t ld_out, uint64_t seed, uint32_t stream_id);
// a pass over the SNP-major matrix in segments (resident rows, then windows re-created from the sample-major matrix)
int for_each_gs_segment(gpca_ctx* c, const std::function<int(const PackedMat&, uint64_t)>& fn);
void gpca_comm_destroy(gpca_ctx* c);   // comm.cu
// generic timed sketch on an arbitrary view (used by the EigenSNP driver)
int timed_sketch(gpca_ctx* c, const SketchProblem& p);
// one launch for all LD blocks (integer engine, item mode)
int timed_sketch_batch(gpca_ctx* c, const SketchBatch& sb);
