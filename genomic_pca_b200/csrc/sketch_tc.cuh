// sketch_tc.cuh -- tcgen05 / TMEM / TMA engine for the sketch pass (engine 1).
#pragma once
#include "kernels.cuh"

// true when the tcgen05 engine can take this problem (shape/alignment); otherwise the SIMT engine runs.
bool sketch_tc_supported(gpca_ctx* c, const SketchProblem& p);
int launch_sketch_tc(gpca_ctx* c, const SketchProblem& p);

// integer engine (engine 2): exact int32 accumulation of u8 codes x two s8 limbs of the dense operand, l <= 32
bool sketch_i8_supported(gpca_ctx* c, const SketchProblem& p);
int launch_sketch_i8(gpca_ctx* c, const SketchProblem& p);

// ---- batched passes over many LD blocks in one launch (integer engine, "item mode") ---------------------------------
// Every LD block has its own dense operand; a work item is (block, group of 256 rows).  Two shapes are used by the
// EigenSNP driver:
//   rows = the block's SNPs (contiguous rows of the matrix), K = all columns         -> out rows are per SNP
//   rows = all samples, K = the block's SNPs (a byte range inside every packed row) -> out is per (block, sample)
struct I8Item {          // 32 bytes, read by the kernel as two 16-byte words
  uint32_t row0;         // first row of the packed matrix
  uint32_t nrows_l;      // valid rows (<= 256) | logical columns << 16
  uint32_t kbyte0;       // first byte of the K range inside a packed row
  uint32_t nst;          // stages of 256 fields
  uint32_t img_st0;      // first stage of the block's operand image
  uint32_t blk;          // block index (column sums, scale)
  uint32_t out_off_lo, out_off_hi;   // element offset of the item's first output row
};
struct SketchBatchBlock {  // operand of one block
  uint64_t bin_off;        // element offset of the block's [K x ld] operand inside Bin
  uint64_t fe_off;         // offset into f / e for the block's first K index
  uint32_t K;              // rows of the operand (= fields of the K range)
  uint32_t l;              // logical columns (<= 32)
  uint32_t img_st0, nst;   // image stages [img_st0, img_st0 + nst), nst = ceil(K / 256)
  uint32_t kskip = 0;      // the first kskip rows count as zero: fields in front of a block whose K range had to start
  uint32_t pad_ = 0;       //   at the 16-byte boundary below the block's first field
};
struct SketchBatch {
  PackedMat G;                       // the whole packed matrix the items index into
  const I8Item* d_items;
  uint32_t n_items;
  const SketchBatchBlock* d_blocks;
  uint32_t n_blocks;
  uint32_t total_img_stages;
  uint32_t max_K;                    // largest operand (sizes the statistics grid)
  const float* Bin;
  uint32_t ld;
  const float* f;                    // per-K scale / rank-one weight, indexed fe_off + k (nullptr = 1)
  const float* e;
  const float* a;                    // per-row scale / rank-one weight, indexed by matrix row (nullptr = 1)
  const float* b;
  float* out;
  uint32_t ldo;
  double bytes;                      // algorithmic bytes of the pass (statistics)
  bool wide_boxes = false;           // 128-byte TMA boxes (long items on a matrix with a large row pitch), else 64-byte
};
bool sketch_i8_batch_supported(gpca_ctx* c);
int launch_sketch_i8_batch(gpca_ctx* c, const SketchBatch& sb);
