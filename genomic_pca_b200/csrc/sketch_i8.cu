// sketch_i8.cu -- integer tensor-core engine for the sketch pass (engine 2, l <= 32).
//
//   acc[r, :] = sum_k code(r,k) * B[k, :]    computed EXACTLY in int32 from
//     A = the 2-bit dosage codes widened to u8 in registers (4 genotypes per register:
//         r_j = (w >> 2j) & 0x03030303, shifts on the FMA pipe, one LOP3 per 4 genotypes) and stored to TMEM,
//     B = the dense operand quantised to 16-bit fixed point (relative to max|f o B|) and split into two
//         signed 8-bit limbs q = 256*hi + lo, hi/lo in [-128,127]  ->  N = 64 columns (hi | lo) per MMA,
//   with tcgen05.mma.kind::i8 (u8 x s8 -> s32, A from TMEM, B from smem).  The epilogue recombines
//   256*D_hi + D_lo (int -> fp32, one FMA), rescales, and applies the rank-one standardisation correction.
//
// Why a second tensor engine: in the fp16 engine the ALU pipe (LOP3, 64 lanes/clk/SM) is the
// co-limiter with HBM -- 9 ALU ops per 16 genotypes; here it is 4, TMEM store traffic is halved, a TMEM slot holds
// twice the genotypes, the accumulation is exact (bit-reproducible for any K split), and the only rounding left is the
// 16-bit quantisation of the dense operand (about 2^-15 of the column maximum, tighter than fp16's 2^-11).
// Same warp roles / rings / persistent scheduling as sketch_tc.cu.
#include <cuda.h>

#include "sketch_tc.cuh"
#include "tc_ptx.cuh"
#include "philox.cuh"

namespace {
using namespace tcptx;

// Row tiles of 128 rows per CTA: a template parameter of the kernel.  RT = 2 (256 rows, 256 TMEM columns, two CTAs
// per SM) is the general configuration.  RT = 4 (512 rows, the whole TMEM, one CTA per SM, 20 warps) shares every B'
// stage between twice as many rows: the operand image comes out of L2 once per 512 rows instead of once per 256, which
// halves the L2 -> SM operand traffic -- as large as the genotype stream itself at RT = 2 (a 16 KB image stage per
// 16 KB of packed rows).  It is used for long items (launch_sketch_i8: the per-item epilogue is not overlapped by a
// second CTA then).
constexpr int STAGE_FIELDS = 256;  // 64 B per row per stage
constexpr int CHUNKS = 4;          // 64-field chunks per stage
// Bytes of a packed row per A stage = TMA box width.  Regular passes use 128 (two 256-field stages per A stage): 64-byte
// boxes cap the TMA stream at ~2.6 TB/s when the row pitch is MBs -- one DRAM page per 64 B -- while 128-byte boxes
// reach ~5.1 TB/s (tools/probe/tma_bw_probe.cu).  Item mode (batched LD blocks: one or two stages per item, bound by
// the latency of an item rather than by bandwidth) keeps 64-byte boxes so that four stages can be in flight.
// BOX = bytes of a packed row per TMA box (64 or 128), see the comment above
#ifndef GPCA_I8_DEEP_KH
#define GPCA_I8_DEEP_KH 1
#endif
// DEEP (RT = 2 only): ONE 256-row CTA per SM that owns the whole TMEM -- six A slots instead of two, so the expanders
// run up to six chunk pairs ahead of the MMAs -- with two expander warps per (row tile, lane quarter), each expanding
// one 64-field half of every chunk pair (16 expander warps per SM, as with two regular CTAs).
template <int BOX, int RT, bool DEEP = false>
struct ACfg {
  static_assert(!DEEP || (RT == 2 && BOX == 128), "the deep-slot shape is a 256-row CTA with 128-byte boxes");
  static constexpr int ROW_BYTES = BOX;
  static constexpr int HALVES = ROW_BYTES / 64;          // 256-field stages per A stage
  static constexpr int TILE_BYTES = 128 * ROW_BYTES;
  static constexpr int STAGE_BYTES = RT * TILE_BYTES;
  static constexpr int RING_BYTES = DEEP ? 131072 : 32768 * RT;   // 64 KB of A stages at RT = 2, 128 KB at RT = 4 / DEEP
  static constexpr int SA = RING_BYTES / STAGE_BYTES;
  static constexpr int KH = (DEEP && GPCA_I8_DEEP_KH == 2) ? 2 : 1;   // expander warps per (row tile, lane quarter): K halves of a chunk pair
  static constexpr int NUM_THREADS = 128 + 128 * RT * KH;   // warps 0..3: producers / issuer / idle; then 4 * KH expander warps per row tile
                                                         // (a multiple of 4 warps: warp & 3 is the TMEM lane quarter of the warp)
  static constexpr int TMEM_COLS = (RT == 2 && !DEEP) ? 256 : 512;
  static constexpr int A_COL0 = RT * 64;                 // accumulators: RT * 64 columns, then SLOTS * RT * 32 columns of A slots
  static constexpr int CTAS_PER_SM = (RT == 2 && !DEEP) ? 2 : 1;
  static constexpr int SLOTS = DEEP ? 6 : 2;             // a TMEM slot holds a chunk pair (128 fields) of every row tile
  static constexpr int SB = DEEP ? 4 : 3;                // B' stages
  static_assert(RT * 64 + SLOTS * RT * 32 <= TMEM_COLS, "TMEM budget");
};
constexpr int NL = 32;             // logical columns
constexpr int NM = 64;             // MMA N = hi | lo limbs
#ifndef GPCA_I8_REGSPLIT
#define GPCA_I8_REGSPLIT 1
#endif
// (RT = 2: 384 threads compiled at 80 registers; warpgroup 0 gives 128 x 48 registers up, the two expander warpgroups
//  take 256 x 24.  RT = 4 runs one CTA per SM at the compiled allocation.)
#if GPCA_I8_REGSPLIT
#define REG_DEC() do { if (RT == 2 && !DEEP) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); } while (0)
#define REG_INC() do { if (RT == 2 && !DEEP) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); } while (0)
#else
#define REG_DEC()
#define REG_INC()
#endif
// GPCA_I8_PROF: per-warp cycle accounting of the waits (printed by CTA 0 at the end of every launch; measurement builds only)
#ifdef GPCA_I8_PROF
#include <cstdio>
#define PROF_DECL(...) uint32_t __VA_ARGS__
#define PROF_T(x) const uint32_t x = prof_clock()
#define PROF_ADD(acc, t0) acc += prof_clock() - (t0)
__device__ __forceinline__ uint32_t prof_clock() { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }
#else
#define PROF_DECL(...)
#define PROF_T(x)
#define PROF_ADD(acc, t0)
#endif
// GPCA_I8_DEFER_ST: tcgen05.wait::st of a chunk pair is issued after the NEXT pair has been expanded (the stores drain
// while the warp computes) instead of right behind the stores.
#ifndef GPCA_I8_DEFER_ST
#define GPCA_I8_DEFER_ST 0
#endif
// GPCA_I8_L2PF = n: the A producer prefetches the stage n ahead into L2 (0 = off)
#ifndef GPCA_I8_L2PF
#define GPCA_I8_L2PF 0
#endif
#ifndef GPCA_I8_PIN_EXPAND
#define GPCA_I8_PIN_EXPAND 0
#endif
// GPCA_I8_STAGGER (RT = 2, regular shape): the two row tiles of a CTA take turns storing a chunk pair to TMEM (tile 0,
// then tile 1), so that at most 8 of the SM's 16 expander warps drive the TMEM store port at a time.  With all 16
// storing at once the port saturates (221 B/clk) and a concurrent tcgen05.mma takes 45 cycles instead of 32; with 8 it
// takes 33.5 (tools/probe/mma_sttm_probe.cu).
#ifndef GPCA_I8_STAGGER
#define GPCA_I8_STAGGER 0
#endif
// GPCA_I8_TRACE (with GPCA_I8_PROF): %clock timestamps of the hand-over events of 96 consecutive chunk pairs, for the two
// CTAs resident on SM 0, printed at the end of the launch (measurement builds only).
#if defined(GPCA_I8_TRACE) && defined(GPCA_I8_PROF)
#define TR_FIRST 64u
#define TR_COUNT 96u
__device__ unsigned g_tr[2][TR_COUNT][8];
__device__ unsigned g_tr_slot;
#define TR(pair, ev) do { if (tr_slot >= 0 && (pair) >= TR_FIRST && (pair) < TR_FIRST + TR_COUNT) g_tr[tr_slot][(pair) - TR_FIRST][ev] = prof_clock(); } while (0)
#else
#define TR(pair, ev)
#endif
#ifndef GPCA_I8_TILE_SYNC_DEFAULT
#define GPCA_I8_TILE_SYNC_DEFAULT false
#endif
constexpr int B_STAGE_BYTES = STAGE_FIELDS * NM;       // 1 byte per element
constexpr int D_COL0 = 0;
template <int RT, bool DEEP = false>
constexpr int smem_bytes_for() { return ACfg<128, RT, DEEP>::RING_BYTES + ACfg<128, RT, DEEP>::SB * B_STAGE_BYTES + 512 + 320; }
constexpr uint32_t MAX_STAGES_PER_ITEM = 20000;        // 3 * 128 * 256 * 20000 < 2^31: no int32 overflow

// 16 fields of a word -> 4 registers of 4 x u8 (register j holds fields j, j+4, j+8, j+12)
__device__ __forceinline__ void expand_word_u8(uint32_t w, uint32_t* r) {
  r[0] = w & 0x03030303u;
  r[1] = __umulhi(w, 1u << 30) & 0x03030303u;   // w >> 2 on the FMA pipe
  r[2] = __umulhi(w, 1u << 28) & 0x03030303u;   // w >> 4
  r[3] = __umulhi(w, 1u << 26) & 0x03030303u;   // w >> 6
}

struct I8Params {
  const int8_t* bimg;   // [total_stages][STAGE_FIELDS * 64] bytes (UMMA K-major core-matrix image, hi|lo limbs)
  uint64_t rows;
  uint32_t total_stages, stages_per_split, ksplit, row_groups, n_items;
  const float* a;
  const float* b;
  const float* cvec;    // [32]            (item mode: [n_blocks][32])
  const float* scales;  // [0] = quantisation scale, [1] = dequantisation scale   (item mode: [n_blocks][2])
  float* out;
  uint32_t ldo, l;
  float* partial;       // [ksplit][rows][32]
  const I8Item* items;  // item mode (batched per-LD-block passes): explicit work list, no split-K partials
  unsigned int* stat_amax;   // regular mode, direct output: max |a_r * out[r,:]| (float bits) as a by-product, or null
  uint32_t pad_cols;         // columns [l, pad_cols) of every output row are padding owned by this pass (SketchProblem::out_pad)
};

// What one work item covers.  Regular mode derives it from (k-split, row group); item mode reads it from the table.
struct ItemInfo {
  uint32_t row0;      // first row of the packed matrix
  uint32_t nrows;     // valid rows (<= 256)
  uint32_t kbyte0;    // first byte (4 fields) of the K range inside a packed row
  uint32_t nst;       // stages of 256 fields
  uint32_t img_st0;   // first stage of the operand image
  uint32_t blk;       // operand block (column sums / scale), item mode
  uint32_t l;         // logical output columns
  uint32_t ks;        // k-split index, regular mode
  uint64_t out_off;   // element offset of the item's first output row
};

template <bool ITEMS, int RT>
__device__ __forceinline__ ItemInfo decode_item(const I8Params& p, uint32_t item) {
  ItemInfo ii;
  if (ITEMS) {
    const uint4* q = reinterpret_cast<const uint4*>(p.items + item);
    const uint4 i0 = __ldg(q), i1 = __ldg(q + 1);
    ii.row0 = i0.x;
    ii.nrows = i0.y & 0xffffu;
    ii.l = i0.y >> 16;
    ii.kbyte0 = i0.z;
    ii.nst = i0.w;
    ii.img_st0 = i1.x;
    ii.blk = i1.y;
    ii.out_off = (uint64_t)i1.z | ((uint64_t)i1.w << 32);
    ii.ks = 0;
  } else {
    const uint32_t ks = item / p.row_groups, rg = item - ks * p.row_groups;
    const uint32_t st0 = ks * p.stages_per_split;
    uint32_t st1 = st0 + p.stages_per_split;
    if (st1 > p.total_stages) st1 = p.total_stages;
    ii.row0 = rg * (RT * 128);
    const uint64_t left = p.rows - (uint64_t)ii.row0;
    ii.nrows = left < (uint64_t)(RT * 128) ? (uint32_t)left : (uint32_t)(RT * 128);
    ii.kbyte0 = st0 * 64;
    ii.nst = st1 - st0;
    ii.img_st0 = st0;
    ii.blk = 0;
    ii.l = p.l;
    ii.ks = ks;
    ii.out_off = (uint64_t)ii.row0 * p.ldo;
  }
  return ii;
}

// TS (tile sync): the TMEM hand-over between expanders and the MMA issuer is per ROW TILE (barriers of 4 warps) instead
// of per CTA (all 4 * RT expander warps): a tile's MMAs start as soon as its own four warps have stored their chunk pair,
// and its warps get the slot back without waiting for the other tile's MMAs.
template <bool ITEMS, int RT, bool TS, int BOX, bool DEEP = false>
__global__ void __launch_bounds__(ACfg<BOX, RT, DEEP>::NUM_THREADS, ACfg<BOX, RT, DEEP>::CTAS_PER_SM)
sketch_i8_kernel(const __grid_constant__ CUtensorMap tmap, const I8Params p) {
  static_assert(!(DEEP && ITEMS), "deep slots: regular mode");
  // DEEP with the per-tile hand-over runs one MMA issuer per row tile (warps 1 and 3): a single thread issues one
  // tcgen05.mma per ~48 cycles at best (tools/probe/mma_sttm_probe.cu) -- two CTAs per SM have two issuers between them,
  // a lone CTA needs two of its own to keep the pipe (32 cycles per MMA at N = 64) busy.
  constexpr bool DI = DEEP && TS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t a_ring = smem_base;
  using AC = ACfg<BOX, RT, DEEP>;
  constexpr int A_ROW_BYTES = AC::ROW_BYTES, HALVES = AC::HALVES, A_TILE_BYTES = AC::TILE_BYTES,
                A_STAGE_BYTES = AC::STAGE_BYTES, SA = AC::SA, A_RING_BYTES = AC::RING_BYTES, TMEM_COLS = AC::TMEM_COLS,
                A_COL0 = AC::A_COL0, SLOTS = AC::SLOTS, SB = AC::SB, KH = AC::KH;
  static_assert((RT == 2 || RT == 4) && SA * A_STAGE_BYTES == A_RING_BYTES && SA >= 2 && SA <= 4, "A ring layout");
  const uint32_t b_ring = smem_base + A_RING_BYTES;
  const uint32_t bars = b_ring + SB * B_STAGE_BYTES;
  auto bar_afull = [&](int s) { return bars + 8u * s; };
  auto bar_aempty = [&](int s) { return bars + 8u * (4 + s); };
  auto bar_bfull = [&](int s) { return bars + 8u * (8 + s); };
  // TMEM slot barriers: one pair per slot, or (TS) one pair per (slot, row tile)
  auto bar_tfull = [&](int j, int t) { return bars + 8u * (32 + (TS ? j * RT + t : j)); };
  auto bar_tempty = [&](int j, int t) { return bars + 8u * (48 + (TS ? j * RT + t : j)); };
  static_assert(SLOTS * (TS ? RT : 1) <= 16, "TMEM slot barriers");
  auto bar_bempty = [&](int s) { return bars + 8u * (24 + s); };
  auto bar_turn = [&](int t) { return bars + 8u * (12 + t); };     // GPCA_I8_STAGGER: "row tile t has stored its chunk pair"
  const uint32_t bar_accfull = bars + 8u * 28;
  const uint32_t bar_accempty = bars + 8u * 29;
  const uint32_t tmem_slot = bars + 8u * 30;
  const uint32_t cvec_smem = bars + 512;
  float* cv_s = reinterpret_cast<float*>(smem_raw + (cvec_smem - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(bar_afull(s), 1);
      mbar_init(bar_aempty(s), 4 * RT * KH);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_bempty(s), DI ? 2 : 1);
    }
    for (int j = 0; j < SLOTS; ++j)
      for (int t = 0; t < (TS ? RT : 1); ++t) {
        mbar_init(bar_tfull(j, t), TS ? 4 * KH : 4 * RT * KH);
        mbar_init(bar_tempty(j, t), 1);
      }
    mbar_init(bar_turn(0), 4);
    mbar_init(bar_turn(1), 4);
    mbar_init(bar_accfull, DI ? 2 : 1);
    mbar_init(bar_accempty, 4 * RT);
    fence_barrier_init();
  }
#if defined(GPCA_I8_TRACE) && defined(GPCA_I8_PROF)
  __shared__ int tr_slot_s;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr_slot_s = (smid == 0 && !ITEMS) ? (int)(atomicAdd(&g_tr_slot, 1u) & 1u) : -1;
  }
#endif
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (!ITEMS) {
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
      const int i = threadIdx.x - 64;
      cv_s[i] = (i < NL) ? p.cvec[i] : 0.0f;
    } else if (threadIdx.x == 128) {
      cv_s[64] = p.scales[1];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const float s_scale = ITEMS ? 0.0f : cv_s[64];
#if defined(GPCA_I8_TRACE) && defined(GPCA_I8_PROF)
  const int tr_slot = tr_slot_s;
#endif
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // warpgroup 0 = the two producers, the MMA issuer and an idle warp; warpgroups 1 and 2 = the expanders of row tile
  // 0 and 1.  Registers move from the first to the others (setmaxnreg works per warpgroup; every warp of the group
  // executes it, at the top of its own role branch).
  if (warp == 0 || warp == 2) {
    REG_DEC();
    // ---- producers: warp 0 = packed genotype tiles (TMA 2-D), warp 2 = B image (1-D bulk copies)
    // ONE elected thread runs the whole role.  (Electing a lane around every issue -- `if (elect_one()) {...}
    // __syncwarp();` per stage / per group of MMAs -- costs ~100 cycles of divergence and reconvergence each time:
    // tools/probe/mma_sttm_probe.cu measures 74 cycles per tcgen05.mma with four MMAs per elected region against 32,
    // the pipe's own rate, once the regions are gone.  That was what held the MMA issuer, and with it the pass.)
    const bool is_a = (warp == 0);
    uint32_t it = 0;
    PROF_DECL(c_wait = 0);
    PROF_T(t_role);
    const bool leader = elect_one();
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x) {
      const ItemInfo ii = decode_item<ITEMS, RT>(p, item);
      const int row0 = (int)ii.row0;
      if (is_a) {
        const uint32_t n_ast = (ii.nst + HALVES - 1) / HALVES;
        for (uint32_t a = 0; a < n_ast; ++a, ++it) {
          const int s = it % SA;
          const uint32_t ph = (it / SA) & 1u;
          PROF_T(t0);
          mbar_wait(bar_aempty(s), ph ^ 1u);
          PROF_ADD(c_wait, t0);
          const uint32_t sbase = a_ring + s * A_STAGE_BYTES;
          mbar_arrive_expect_tx(bar_afull(s), A_STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < RT; ++t)
            tma_load_2d(sbase + t * A_TILE_BYTES, &tmap, bar_afull(s), (int)(ii.kbyte0 + a * A_ROW_BYTES),
                        row0 + t * 128);
#if GPCA_I8_L2PF
          // The ring holds SA stages; the stage SA ahead is requested only when this one has been consumed -- about one
          // HBM round trip before it is needed (trace: every new A stage arrived ~400 cycles late).  Pulling it into L2
          // now makes that later load an L2 hit.
          if (!ITEMS && a + GPCA_I8_L2PF < n_ast) {
#pragma unroll
            for (int t = 0; t < RT; ++t)
              tma_prefetch_2d(&tmap, (int)(ii.kbyte0 + (a + GPCA_I8_L2PF) * A_ROW_BYTES), row0 + t * 128);
          }
#endif
        }
      } else {
        for (uint32_t st = 0; st < ii.nst; ++st, ++it) {
          const int s = it % SB;
          const uint32_t ph = (it / SB) & 1u;
          PROF_T(t0);
          mbar_wait(bar_bempty(s), ph ^ 1u);
          PROF_ADD(c_wait, t0);
#ifdef GPCA_KO_BLOAD      // (measurement build: the image is copied for the first SB stages only -- wrong results, timing only)
          if (it >= (uint32_t)SB) {
            mbar_arrive(bar_bfull(s));
            continue;
          }
#endif
          mbar_arrive_expect_tx(bar_bfull(s), B_STAGE_BYTES);
          bulk_load_1d(b_ring + s * B_STAGE_BYTES, p.bimg + (size_t)(ii.img_st0 + st) * B_STAGE_BYTES, B_STAGE_BYTES,
                       bar_bfull(s));
        }
      }
    }
#ifdef GPCA_I8_PROF
    if (blockIdx.x == 0 && leader)
      printf("PROF %s producer: total %u wait_empty %u\n", is_a ? "A" : "B", prof_clock() - t_role, c_wait);
#endif
    __syncwarp();
  } else if (warp == 1 || (DI && warp == 3)) {
    REG_DEC();
    const int t_lo = DI ? (warp == 1 ? 0 : 1) : 0, t_hi = DI ? t_lo + 1 : RT;
    // ---- MMA issuer: D = s32, A = u8 (TMEM), B = s8 (smem, K-major, no swizzle), M = 128, N = 64, K = 32
    const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NM >> 3) << 17) | (8u << 24);
    // B descriptor: LBO = 64 rows * 16 B = 1024 B (next 16-wide K chunk), SBO = 128 B (next 8 columns), version 1
    const uint32_t desc_lo_const = (uint32_t)(1024 >> 4) << 16;
    const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);
    uint32_t it = 0, cit = 0, item_idx = 0;
    PROF_DECL(c_acc = 0, c_bfull = 0, c_tfull = 0);
    PROF_T(t_role);
    // one elected thread runs the whole role (see the producers above): waits, MMAs and commits without a divergent
    // region per group of MMAs
    const bool leader = elect_one();
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x, ++item_idx) {
      const uint32_t nst = decode_item<ITEMS, RT>(p, item).nst;
      PROF_T(t_a);
      mbar_wait(bar_accempty, (item_idx & 1u) ^ 1u);
      PROF_ADD(c_acc, t_a);
      tc_fence_after();
      uint32_t acc_flag = 0;
      for (uint32_t st = 0; st < nst; ++st, ++it) {
        const int s = it % SB;
        const uint32_t ph = (it / SB) & 1u;
        PROF_T(t_b);
        mbar_wait(bar_bfull(s), ph);
        PROF_ADD(c_bfull, t_b);
        const uint32_t bsm = b_ring + s * B_STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < CHUNKS; q += 2, ++cit) {
          const int slot = cit % SLOTS;
          const uint32_t sph = (cit / SLOTS) & 1u;
          if (!TS) {
            PROF_T(t_t);
            mbar_wait(bar_tfull(slot, 0), sph);
            PROF_ADD(c_tfull, t_t);
            tc_fence_after();
            TR(cit, 4);
          }
#pragma unroll
          for (int t = 0; t < RT; ++t) {
            if (t < t_lo || t >= t_hi) continue;
            if (TS) {
              PROF_T(t_t);
              mbar_wait(bar_tfull(slot, t), sph);
              PROF_ADD(c_tfull, t_t);
              tc_fence_after();
            }
            const uint32_t d_t = tmem_base + D_COL0 + t * NM;
            const uint32_t a_t = tmem_base + A_COL0 + (slot * RT + t) * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {      // 4 MMAs of K = 32 cover the 128 fields of the pair
              const uint32_t baddr = bsm + (uint32_t)((q * 2 + i) * (32 * NM));
              const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo_const | ((baddr >> 4) & 0x3FFFu));
#ifndef GPCA_KO_MMA
              tc_mma_ts_i8(d_t, a_t + 8 * i, bdesc, idesc, acc_flag | (uint32_t)i);
#else
              if (i == 0 && GPCA_KO_MMA) tc_mma_ts_i8(d_t, a_t + 8 * i, bdesc, idesc, acc_flag | (uint32_t)i);
#endif
            }
            if (TS || t == t_hi - 1) tc_commit(bar_tempty(slot, TS ? t : 0));
          }
          TR(cit, 5);
          acc_flag = 1;
        }
        tc_commit(bar_bempty(s));
      }
      tc_commit(bar_accfull);
    }
#ifdef GPCA_I8_PROF
    if (blockIdx.x == 0 && leader)
      printf("PROF issuer: total %u wait_accempty %u wait_bfull %u wait_tfull %u pairs %u\n", prof_clock() - t_role, c_acc,
             c_bfull, c_tfull, cit);
#endif
    __syncwarp();
  } else if (warp == 3) {
    REG_DEC();
  } else {
    REG_INC();
    // ---- expanders + epilogue
    const int tile = ((warp - 4) >> 2) % RT;
    const int khalf = KH == 2 ? (warp - 4) / (4 * RT) : 0;      // DEEP: which 64-field half of every chunk pair this warp expands
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // 16-byte chunk c of row r sits at chunk c ^ (r & 7) under SWIZZLE_128B (128-byte rows) and at c ^ ((r >> 1) & 3)
    // under SWIZZLE_64B (64-byte rows)
    const uint32_t sw = (A_ROW_BYTES == 128) ? (uint32_t)(row_in_tile & 7) : (uint32_t)((row_in_tile >> 1) & 3);
    uint32_t it = 0, cit = 0, item_idx = 0;
    float stat_max = 0.0f;     // by-product statistic of the output (see SketchProblem::emit_stats)
#if GPCA_I8_DEFER_ST
    bool st_pending = false;
    int pend_slot = 0;
#endif
    PROF_DECL(c_afull = 0, c_tempty = 0, c_st = 0, c_epi = 0, c_accfull = 0, c_ldtm = 0, c_ab = 0);
    PROF_T(t_role);
    for (uint32_t item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_idx) {
      const ItemInfo ii = decode_item<ITEMS, RT>(p, item);
      const uint32_t n_ast = (ii.nst + HALVES - 1) / HALVES;
      for (uint32_t a = 0; a < n_ast; ++a, ++it) {
        const int s = it % SA;
        const uint32_t ph = (it / SA) & 1u;
        PROF_T(t_af);
        mbar_wait(bar_afull(s), ph);
        PROF_ADD(c_afull, t_af);
        const uint32_t arow = a_ring + s * A_STAGE_BYTES + tile * A_TILE_BYTES + row_in_tile * A_ROW_BYTES;
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          if (a * HALVES + h >= ii.nst) break;     // odd stage count: the last A stage is half used
          if (DEEP && KH == 2) {
            // one 16-byte chunk (64 fields) of each of the two chunk pairs of the stage: 16 columns of the slot
            uint4 v[CHUNKS / 2];
#pragma unroll
            for (int q = 0; q < CHUNKS / 2; ++q) {
              const uint32_t addr = arow + (((uint32_t)(h * CHUNKS + 2 * q + khalf) ^ sw) << 4);
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w)
                           : "r"(addr));
            }
#pragma unroll
            for (int q = 0; q < CHUNKS / 2; ++q, ++cit) {
              const int slot = cit % SLOTS;
              const uint32_t sph = (cit / SLOTS) & 1u;
              uint32_t r0[16];
              expand_word_u8(v[q].x, r0 + 0);
              expand_word_u8(v[q].y, r0 + 4);
              expand_word_u8(v[q].z, r0 + 8);
              expand_word_u8(v[q].w, r0 + 12);
              PROF_T(t_te);
              mbar_wait(bar_tempty(slot, tile), sph ^ 1u);
              PROF_ADD(c_tempty, t_te);
              tc_fence_after();
              PROF_T(t_st);
              tmem_st16(tmem_base + lane_addr + A_COL0 + (slot * RT + tile) * 32 + 16 * khalf, r0);
              tc_wait_st();
              PROF_ADD(c_st, t_st);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tfull(slot, tile));
            }
            continue;
          }
          uint4 v[CHUNKS];
#pragma unroll
          for (int q = 0; q < CHUNKS; ++q) {
            const uint32_t addr = arow + (((uint32_t)(h * CHUNKS + q) ^ sw) << 4);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w)
                         : "r"(addr));
          }
#pragma unroll
          for (int q = 0; q < CHUNKS; q += 2, ++cit) {
            const int slot = cit % SLOTS;
            const uint32_t sph = (cit / SLOTS) & 1u;
            uint32_t r0[16], r1[16];
#ifdef GPCA_KO_EXPAND
#pragma unroll
            for (int z = 0; z < 16; ++z) { r0[z] = (&v[q].x)[z & 3] + z; r1[z] = (&v[q + 1].x)[z & 3] + z; }
#else
            expand_word_u8(v[q].x, r0 + 0);
            expand_word_u8(v[q].y, r0 + 4);
            expand_word_u8(v[q].z, r0 + 8);
            expand_word_u8(v[q].w, r0 + 12);
            expand_word_u8(v[q + 1].x, r1 + 0);
            expand_word_u8(v[q + 1].y, r1 + 4);
            expand_word_u8(v[q + 1].z, r1 + 8);
            expand_word_u8(v[q + 1].w, r1 + 12);
#endif
#if GPCA_I8_PIN_EXPAND
            // The expansions are plain register arithmetic: without this, the compiler sinks them below the barrier waits
            // that follow (asm volatile), right in front of the tcgen05.st that consumes them -- and the warp then waits
            // for the slot first and computes afterwards.  Naming the results as inputs of an (empty) volatile asm keeps
            // the computation in front of the waits.
            asm volatile("" ::"r"(r0[0]), "r"(r0[1]), "r"(r0[2]), "r"(r0[3]), "r"(r0[4]), "r"(r0[5]), "r"(r0[6]), "r"(r0[7]),
                         "r"(r0[8]), "r"(r0[9]), "r"(r0[10]), "r"(r0[11]), "r"(r0[12]), "r"(r0[13]), "r"(r0[14]), "r"(r0[15]));
            asm volatile("" ::"r"(r1[0]), "r"(r1[1]), "r"(r1[2]), "r"(r1[3]), "r"(r1[4]), "r"(r1[5]), "r"(r1[6]), "r"(r1[7]),
                         "r"(r1[8]), "r"(r1[9]), "r"(r1[10]), "r"(r1[11]), "r"(r1[12]), "r"(r1[13]), "r"(r1[14]), "r"(r1[15]));
#endif
#if GPCA_I8_DEFER_ST
            // the previous pair's TMEM stores were left in flight while this pair was expanded: complete and publish them
            if (st_pending) {
              PROF_T(t_st0);
              tc_wait_st();
              PROF_ADD(c_st, t_st0);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tfull(pend_slot, tile));
              st_pending = false;
            }
#endif
            PROF_T(t_te);
            if (lane == 0 && quarter == 0 && tile == 0) TR(cit, 0);
            mbar_wait(bar_tempty(slot, tile), sph ^ 1u);      // the MMAs that read this slot (of this tile: TS) have completed
            PROF_ADD(c_tempty, t_te);
            if (lane == 0 && quarter == 0 && tile == 0) TR(cit, 1);
            tc_fence_after();
#if GPCA_I8_STAGGER
            if (RT == 2 && !DEEP) {
              if (tile == 0) {
                if (cit > 0) mbar_wait(bar_turn(1), (cit - 1) & 1u);     // tile 1 has stored the previous pair
              } else {
                mbar_wait(bar_turn(0), cit & 1u);                        // tile 0 has stored this pair
              }
            }
#endif
            PROF_T(t_st);
            const uint32_t ta = tmem_base + lane_addr + A_COL0 + (slot * RT + tile) * 32;
#ifndef GPCA_KO_STTM
            tmem_st16(ta, r0);
            tmem_st16(ta + 16, r1);
#if GPCA_I8_DEFER_ST
            st_pending = true;
            pend_slot = slot;
            continue;
#endif
            tc_wait_st();
            PROF_ADD(c_st, t_st);
            if (lane == 0 && quarter == 0) TR(cit, 2 + tile);
            if (lane == 0 && quarter == 3) TR(cit, 6 + tile);
#if GPCA_I8_STAGGER
            if (RT == 2 && !DEEP && lane == 0) mbar_arrive(bar_turn(tile));
#endif
#else
            asm volatile("" :: "r"(r0[0] ^ r0[1] ^ r0[2] ^ r0[3] ^ r0[4] ^ r0[5] ^ r0[6] ^ r0[7] ^ r0[8] ^ r0[9] ^ r0[10] ^ r0[11] ^ r0[12] ^ r0[13] ^ r0[14] ^ r0[15] ^ r1[0] ^ r1[1] ^ r1[2] ^ r1[3] ^ r1[4] ^ r1[5] ^ r1[6] ^ r1[7] ^ r1[8] ^ r1[9] ^ r1[10] ^ r1[11] ^ r1[12] ^ r1[13] ^ r1[14] ^ r1[15]), "r"(ta));
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfull(slot, tile));
          }
        }
        // The stage is released only now: every register loaded from it has been consumed by the expansions above, so
        // all of this warp's shared-memory reads have completed before the producer may let TMA overwrite the stage.
        // (Releasing right after ISSUING the loads let the refill race with loads still in flight when two CTAs
        // share an SM.)
        if (lane == 0) mbar_arrive(bar_aempty(s));
      }
#if GPCA_I8_DEFER_ST
      if (st_pending) {
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfull(pend_slot, tile));
        st_pending = false;
      }
#endif
      // ---- epilogue (DEEP: by the warps of K half 0; the others go on to the next item's stages)
      if (KH == 2 && khalf != 0) continue;
      PROF_T(t_epi);
      const uint32_t lrow = (uint32_t)(tile * 128 + row_in_tile);
      const bool live = lrow < ii.nrows;
      const uint64_t r = (uint64_t)ii.row0 + lrow;
      float ar = 1.0f, br = 1.0f;
      if (!p.partial && live) {
        if (p.a) ar = __ldg(p.a + r);
        if (p.b) br = __ldg(p.b + r);
      }
      PROF_T(t_acf);
      mbar_wait(bar_accfull, item_idx & 1u);
      PROF_ADD(c_accfull, t_acf);
      tc_fence_after();
      uint32_t hi[32], lo[32];
      PROF_T(t_ld);
      tmem_ld32(tmem_base + lane_addr + D_COL0 + tile * NM, hi);
      tmem_ld32(tmem_base + lane_addr + D_COL0 + tile * NM + 32, lo);
      tc_wait_ld();
      PROF_ADD(c_ldtm, t_ld);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty);
#ifdef GPCA_I8_PROF
      {   // time until the per-row factors (loaded before the accumulator wait) are in registers
        PROF_T(t_ab);
        asm volatile("" :: "f"(ar), "f"(br));
        float dummy = ar + br;
        asm volatile("" : "+f"(dummy));
        PROF_ADD(c_ab, t_ab);
      }
#endif
      if (live) {
        // 256 * D_hi + D_lo, rescaled: two int -> fp32 conversions and one FMA per element.  (An earlier version formed
        // the 40-bit integer and went through f64 -- 3 quarter-rate conversions/DMULs per element, a third of the
        // expander warps' time when an item is only 10 stages long.)  The fp32 rounding (2^-24 relative) is far below
        // the 2^-15 quantisation of the operand and the expression is fixed, so results stay bit-reproducible.
        const float sc_lo = ITEMS ? __ldg(p.scales + 2 * ii.blk + 1) : s_scale;
        const float sc_hi = 256.0f * sc_lo;
        if (!ITEMS && p.partial) {
          float* dst = p.partial + ((uint64_t)ii.ks * p.rows + r) * NL;
#pragma unroll
          for (int c = 0; c < NL; c += 4) {
            float4 o;
            o.x = fmaf((float)(int)hi[c + 0], sc_hi, (float)(int)lo[c + 0] * sc_lo);
            o.y = fmaf((float)(int)hi[c + 1], sc_hi, (float)(int)lo[c + 1] * sc_lo);
            o.z = fmaf((float)(int)hi[c + 2], sc_hi, (float)(int)lo[c + 2] * sc_lo);
            o.w = fmaf((float)(int)hi[c + 3], sc_hi, (float)(int)lo[c + 3] * sc_lo);
            *reinterpret_cast<float4*>(dst + c) = o;
          }
        } else {
          float* dst = p.out + ii.out_off + (uint64_t)lrow * p.ldo;
          const float* cv = ITEMS ? p.cvec + (size_t)ii.blk * NL : cv_s;
          // widest aligned store the row stride allows.  A thread owns a whole output row, so every store instruction
          // of a warp touches 32 different lines: with 32-byte stores (st.global.v8, sm_100) each one writes a full
          // sector and a padded 32-float row takes 4 instructions (the pad columns inside the row stride get zeros).
          const uintptr_t obase = reinterpret_cast<uintptr_t>(p.out + ii.out_off);
#ifdef GPCA_I8_NO_V8
          const bool vec8 = false;
#else
          const bool vec8 = (p.ldo & 7u) == 0 && (obase & 31) == 0;
#endif
          const bool vec4 = (p.ldo & 3u) == 0 && (obase & 15) == 0;
          const bool vec2 = (p.ldo & 1u) == 0 && (obase & 7) == 0;
#pragma unroll
          for (int c = 0; c < NL; c += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float cj = ITEMS ? __ldg(cv + c + j) : cv[c + j];
              v[j] = ar * fmaf((float)(int)hi[c + j], sc_hi, (float)(int)lo[c + j] * sc_lo) - br * cj;
              if ((uint32_t)(c + j) >= ii.l) v[j] = 0.0f;
              hi[c + j] = __float_as_uint(v[j]);       // kept for the statistics below
            }
            // (a 32-byte store may run past column l only into padding that belongs to this output: with a column view
            //  into a wider matrix -- the condensed features of EigenSNP -- the neighbours' columns live there)
            if (vec8 && ((uint32_t)c + 8 <= ii.l || ((uint32_t)c < ii.l && (uint32_t)c + 8 <= p.pad_cols))) {
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c), "f"(v[0]), "f"(v[1]),
                           "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                           : "memory");
            } else {
#pragma unroll
              for (int h = 0; h < 8; h += 4) {
                if (vec4 && (uint32_t)(c + h) + 3 < ii.l) {
                  *reinterpret_cast<float4*>(dst + c + h) = make_float4(v[h], v[h + 1], v[h + 2], v[h + 3]);
                } else {
#pragma unroll
                  for (int j = h; j < h + 4; j += 2) {
                    if (vec2 && (uint32_t)(c + j) + 1 < ii.l) {
                      *reinterpret_cast<float2*>(dst + c + j) = make_float2(v[j], v[j + 1]);
                    } else {
                      if ((uint32_t)(c + j) < ii.l) dst[c + j] = v[j];
                      if ((uint32_t)(c + j) + 1 < ii.l) dst[c + j + 1] = v[j + 1];
                    }
                  }
                }
              }
            }
          }
        }
      }
      if (!ITEMS && p.stat_amax && live) {
        // by-product statistic of the row just written: max |a_r * out[r,:]| (what the next pass's quantisation needs)
#pragma unroll
        for (int c = 0; c < NL; ++c)
          if ((uint32_t)c < ii.l) stat_max = fmaxf(stat_max, fabsf(ar * __uint_as_float(hi[c])));
      }
      PROF_ADD(c_epi, t_epi);
    }
#ifdef GPCA_I8_PROF
    if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0)
      printf("PROF expander w%d: total %u wait_afull %u wait_tempty %u st_to_waitst %u epilogue %u (wait_accfull %u ldtm %u ab %u) pairs %u items %u\n",
             warp, prof_clock() - t_role, c_afull, c_tempty, c_st, c_epi, c_accfull, c_ldtm, c_ab, cit, item_idx);
#endif
    if (!ITEMS && p.stat_amax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_max = fmaxf(stat_max, __shfl_xor_sync(0xffffffffu, stat_max, o));
      if (lane == 0 && stat_max > 0.0f && isfinite(stat_max)) atomicMax(p.stat_amax, __float_as_uint(stat_max));
    }
  }
  tc_fence_before();
  __syncthreads();
#if defined(GPCA_I8_TRACE) && defined(GPCA_I8_PROF)
  if (threadIdx.x == 0 && tr_slot >= 0) {
    __threadfence();
    for (unsigned i = 0; i < TR_COUNT; ++i)
      printf("PROF TR %d %u exp_done %u slot_free %u stored_t0 %u stored_t1 %u stored_t0q3 %u stored_t1q3 %u tfull_seen %u mma_issued %u\n",
             tr_slot, TR_FIRST + i, g_tr[tr_slot][i][0], g_tr[tr_slot][i][1], g_tr[tr_slot][i][2], g_tr[tr_slot][i][3],
             g_tr[tr_slot][i][6], g_tr[tr_slot][i][7], g_tr[tr_slot][i][4], g_tr[tr_slot][i][5]);
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- operand preparation ----------------------------------------------------------------------------------------
// max |f o Bin| (skipped when the producer of Bin already left it behind: SketchProblem::emit_stats)
__global__ void __launch_bounds__(256) i8_amax_kernel(const float* __restrict__ bin, uint64_t K, uint32_t l, uint32_t ld,
                                                      const float* __restrict__ f, unsigned int* __restrict__ amax_bits) {
  const int cidx = threadIdx.x & 31;
  const int rr = threadIdx.x >> 5;
  const bool c0 = (uint32_t)cidx < l;
  float mx = 0.0f;
  const uint64_t stride = (uint64_t)gridDim.x * 8;
  for (uint64_t k = (uint64_t)blockIdx.x * 8 + rr; k < K; k += 4 * stride) {
    float x0[4], fk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t kk = k + u * stride;
      const bool live = kk < K;
      x0[u] = (live && c0) ? bin[kk * ld + cidx] : 0.0f;
      fk[u] = (live && f) ? f[kk] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) mx = fmaxf(mx, fabsf(x0[u] * fk[u]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (cidx == 0 && mx > 0.0f && isfinite(mx)) atomicMax(amax_bits, __float_as_uint(mx));
}

__global__ void i8_scale_kernel(unsigned int* __restrict__ amax_bits, float* __restrict__ scales) {
  const float m = __uint_as_float(*amax_bits);
  scales[0] = (m > 0.0f) ? 32512.0f / m : 0.0f;     // |q| <= 32512 = 127*256: both limbs fit s8
  scales[1] = (m > 0.0f) ? m / 32512.0f : 0.0f;
  *amax_bits = 0u;
}

// cvec = sum of the per-CTA partial column sums that prep_b_i8_kernel leaves (fixed order: deterministic)
__global__ void __launch_bounds__(256) i8_cvec_kernel(const double* __restrict__ cpart, int nparts,
                                                      float* __restrict__ cvec) {
  __shared__ double red[8][32];
  const int cidx = threadIdx.x & 31, grp = threadIdx.x >> 5;
  double s = 0.0;
  for (int q = grp; q < nparts; q += 8) s += cpart[(uint64_t)q * 32 + cidx];
  red[grp][cidx] = s;
  __syncthreads();
  if (grp == 0) {
    double t = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][cidx];
    cvec[cidx] = (float)t;
  }
}

// B image: MMA group g covers fields 32g .. 32g+31; K-slot s (0..31) holds field 32g + 16*(c/4) + (c%4) + 4*b with
// c = s/4, b = s%4 (the order in which expand_word_u8 lays the fields into TMEM columns/bytes).
// Element (slot s, column n in 0..63; n < 32: hi limb of logical column n, n >= 32: lo limb of column n-32) lives at
//   g*32*64 + (s/16)*(64*16) + (n/8)*128 + (n%8)*16 + (s%16)       (UMMA K-major core matrices, no swizzle)
// The column sums e^T Bin come along for free: a thread always meets the same column n (strides are multiples of 32),
// sums e_k * Bin[k, n] over its chunks in f64, and the CTA leaves one partial per column.
__global__ void __launch_bounds__(256) prep_b_i8_kernel(const float* __restrict__ bin, uint64_t K, uint64_t Kpad,
                                                        uint32_t l, uint32_t ld, const float* __restrict__ f,
                                                        const float* __restrict__ e, const float* __restrict__ scales,
                                                        int8_t* __restrict__ img, double* __restrict__ cpart) {
  __shared__ double sred[8][NL];
  const uint64_t n_chunks = Kpad / 16;       // one warp per 16-slot K chunk, lane = logical column
  const float qs = scales[0];
  const uint32_t n = threadIdx.x & 31u;
  const bool col_live = n < l;
  // per-k scale / weight vectors can be fetched 16 bytes at a time when they are 16-byte aligned (chunks start at
  // multiples of 16 rows)
  const bool vec_ok = (!f || (reinterpret_cast<uintptr_t>(f) & 15) == 0) && (!e || (reinterpret_cast<uintptr_t>(e) & 15) == 0);
  double csum = 0.0;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t kc = warp0; kc < n_chunks; kc += nwarps) {
    const uint64_t kb = kc * 16;               // the chunk's 16 operand rows are kb .. kb+15, in the order below
    uint32_t whi[4] = {0, 0, 0, 0}, wlo[4] = {0, 0, 0, 0};
    if (kb < K && col_live) {
      const float* src = bin + kb * ld + n;
      float fq[16], ek[16];
      if (kb + 16 <= K && vec_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 fv = f ? __ldg(reinterpret_cast<const float4*>(f + kb) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float4 ev = e ? __ldg(reinterpret_cast<const float4*>(e + kb) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          fq[4 * i + 0] = fv.x * qs; fq[4 * i + 1] = fv.y * qs; fq[4 * i + 2] = fv.z * qs; fq[4 * i + 3] = fv.w * qs;
          ek[4 * i + 0] = ev.x; ek[4 * i + 1] = ev.y; ek[4 * i + 2] = ev.z; ek[4 * i + 3] = ev.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool live = kb + i < K;
          fq[i] = live ? (f ? f[kb + i] : 1.0f) * qs : 0.0f;
          ek[i] = live ? (e ? e[kb + i] : 1.0f) : 0.0f;
        }
      }
      float part = 0.0f;                       // fp32 inside the chunk, f64 across chunks
#pragma unroll
      for (int ss = 0; ss < 16; ++ss) {
        // K-slot ss of the chunk holds operand row (ss >> 2) + 4 * (ss & 3) -- the order in which expand_word_u8 lays
        // the fields into the TMEM columns / bytes
        const int kl = (ss >> 2) + 4 * (ss & 3);
        const float v = (kb + kl < K) ? src[(size_t)kl * ld] : 0.0f;
        part = fmaf(v, ek[kl], part);
        int q = __float2int_rn(v * fq[kl]);
        q = max(-32512, min(32512, q));
        const int h = (q + 128) >> 8;          // floor((q+128)/256): lo = q - 256 h in [-128, 127]
        const int lo = q - 256 * h;
        whi[ss >> 2] |= (uint32_t)(h & 0xff) << (8 * (ss & 3));
        wlo[ss >> 2] |= (uint32_t)(lo & 0xff) << (8 * (ss & 3));
      }
      csum += (double)part;
    }
    const uint64_t g = kc >> 1;
    const uint32_t half_idx = (uint32_t)(kc & 1);
    const uint64_t base = g * 32 * NM + (uint64_t)half_idx * (NM * 16);
    *reinterpret_cast<uint4*>(img + base + (n >> 3) * 128 + (n & 7) * 16) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
    *reinterpret_cast<uint4*>(img + base + ((n + 32) >> 3) * 128 + (n & 7) * 16) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
  }
  sred[threadIdx.x >> 5][threadIdx.x & 31] = csum;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[w][threadIdx.x];
    cpart[(uint64_t)blockIdx.x * NL + threadIdx.x] = t;
  }
}

// The same image straight from the generator (SketchProblem::gen): a warp owns a 16-row K chunk; lane L draws the four
// Philox calls of row L/2, column groups 4(L%2) .. 4(L%2)+3 (every normal is generated exactly once), the 16 x 32
// values cross the warp through shared memory, and lane n then quantises column n of the 16 rows as above.
__global__ void __launch_bounds__(256) prep_b_i8_gauss_kernel(uint64_t seed, uint32_t stream, uint64_t row0, uint64_t K,
                                                              uint64_t Kpad, uint32_t l, const float* __restrict__ f,
                                                              const float* __restrict__ e,
                                                              const float* __restrict__ scales,
                                                              int8_t* __restrict__ img, double* __restrict__ cpart) {
  __shared__ double sred[8][NL];
  __shared__ float xch[8][16][33];
  const uint64_t n_chunks = Kpad / 16;
  const float qs = scales[0];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t n = lane;
  const bool col_live = n < l;
  double csum = 0.0;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t kc = warp0; kc < n_chunks; kc += nwarps) {
    const uint64_t kb = kc * 16;
    uint32_t whi[4] = {0, 0, 0, 0}, wlo[4] = {0, 0, 0, 0};
    if (kb < K) {      // warp-uniform
      const uint32_t gr = lane >> 1;                      // row of the chunk this lane generates
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t cg = (lane & 1u) * 4 + i;
        float z[4];
        philox_normal4(seed, stream, row0 + kb + gr, cg, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) xch[wib][gr][cg * 4 + j] = z[j];
      }
      __syncwarp();
      if (col_live) {
        float part = 0.0f;
#pragma unroll
        for (int ss = 0; ss < 16; ++ss) {
          const int kl = (ss >> 2) + 4 * (ss & 3);        // K-slot order of the expansion (see prep_b_i8_kernel)
          const bool live = kb + kl < K;
          const float v = live ? xch[wib][kl][n] : 0.0f;
          const float fk = live ? (f ? __ldg(f + kb + kl) : 1.0f) : 0.0f;
          const float ek = live ? (e ? __ldg(e + kb + kl) : 1.0f) : 0.0f;
          part = fmaf(v, ek, part);
          int q = __float2int_rn(v * (fk * qs));
          q = max(-32512, min(32512, q));
          const int h = (q + 128) >> 8;
          const int lo = q - 256 * h;
          whi[ss >> 2] |= (uint32_t)(h & 0xff) << (8 * (ss & 3));
          wlo[ss >> 2] |= (uint32_t)(lo & 0xff) << (8 * (ss & 3));
        }
        csum += (double)part;
      }
    }
    const uint64_t g = kc >> 1;
    const uint32_t half_idx = (uint32_t)(kc & 1);
    const uint64_t base = g * 32 * NM + (uint64_t)half_idx * (NM * 16);
    *reinterpret_cast<uint4*>(img + base + (n >> 3) * 128 + (n & 7) * 16) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
    *reinterpret_cast<uint4*>(img + base + ((n + 32) >> 3) * 128 + (n & 7) * 16) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
  }
  sred[threadIdx.x >> 5][threadIdx.x & 31] = csum;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[w][threadIdx.x];
    cpart[(uint64_t)blockIdx.x * NL + threadIdx.x] = t;
  }
}

// a-priori quantisation scale of a generated operand (amax >= max |f o Bin| by construction)
__global__ void i8_set_scale_kernel(float amax, float* __restrict__ scales) {
  scales[0] = (amax > 0.0f) ? 32512.0f / amax : 0.0f;
  scales[1] = (amax > 0.0f) ? amax / 32512.0f : 0.0f;
}

__global__ void __launch_bounds__(256) sketch_reduce_i8_kernel(const float* __restrict__ partial, int nsplit,
                                                               uint64_t rows, const float* __restrict__ a,
                                                               const float* __restrict__ b,
                                                               const float* __restrict__ cvec, float* __restrict__ out,
                                                               uint32_t ldo, uint32_t l,
                                                               unsigned int* __restrict__ stat_amax) {
  const uint64_t total = rows * NL;
  float smax = 0.0f;     // by-product statistic of the output: max |a_r * out[r,:]|
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / NL;
    const uint32_t cc = (uint32_t)(t % NL);
    if (cc >= l) continue;
    float s = 0.0f;
    for (int q = 0; q < nsplit; ++q) s += partial[((uint64_t)q * rows + r) * NL + cc];
    const float ar = a ? a[r] : 1.0f, br = b ? b[r] : 1.0f;
    const float v = ar * s - br * cvec[cc];
    out[r * ldo + cc] = v;
    smax = fmaxf(smax, fabsf(ar * v));
  }
  if (stat_amax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    if ((threadIdx.x & 31) == 0 && smax > 0.0f && isfinite(smax)) atomicMax(stat_amax, __float_as_uint(smax));
  }
}

// ---- batched operand preparation (one launch for all LD blocks) ----------------------------------------------------
// grid = (parts, blocks): column sums (f64 partials, summed in a fixed order afterwards) and max |f o Bin| per block
__global__ void __launch_bounds__(256) i8_colstats_batch_kernel(const float* __restrict__ bin, uint32_t ld,
                                                                const SketchBatchBlock* __restrict__ blocks,
                                                                const float* __restrict__ f,
                                                                const float* __restrict__ e,
                                                                double* __restrict__ cpart,
                                                                unsigned int* __restrict__ amax_bits) {
  __shared__ double red[256];
  __shared__ float redm[256];
  const SketchBatchBlock bk = blocks[blockIdx.y];
  const float* src = bin + bk.bin_off;
  const float* fb = f ? f + bk.fe_off : nullptr;
  const float* eb = e ? e + bk.fe_off : nullptr;
  const int cidx = threadIdx.x & 31;
  const int rr = threadIdx.x >> 5;
  const bool c0 = (uint32_t)cidx < bk.l;
  const uint64_t K = bk.K;
  double acc0 = 0.0;
  float mx = 0.0f;
  const uint64_t stride = (uint64_t)gridDim.x * 8;
  for (uint64_t k = (uint64_t)blockIdx.x * 8 + rr; k < K; k += 4 * stride) {
    float x0[4], ek[4], fk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t kk = k + u * stride;
      const bool live = kk < K && kk >= bk.kskip;
      x0[u] = (live && c0) ? src[kk * ld + cidx] : 0.0f;
      ek[u] = (live && eb) ? eb[kk] : 1.0f;
      fk[u] = (live && fb) ? fb[kk] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc0 += (double)(x0[u] * ek[u]);
      mx = fmaxf(mx, fabsf(x0[u] * fk[u]));
    }
  }
  red[threadIdx.x] = acc0;
  redm[threadIdx.x] = mx;
  __syncthreads();
  if (rr == 0) {
    double s0 = 0.0;
    float m = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      s0 += red[q * 32 + cidx];
      m = fmaxf(m, redm[q * 32 + cidx]);
    }
    cpart[((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 + cidx] = s0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (cidx == 0 && m > 0.0f && isfinite(m)) atomicMax(amax_bits + blockIdx.y, __float_as_uint(m));
  }
}

// one warp per block
__global__ void i8_finalize_stats_batch_kernel(const double* __restrict__ cpart, int nparts, uint32_t n_blocks,
                                               float* __restrict__ cvec, unsigned int* __restrict__ amax_bits,
                                               float* __restrict__ scales) {
  const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int cidx = threadIdx.x & 31;
  if (b >= n_blocks) return;
  double s = 0.0;
  for (int q = 0; q < nparts; ++q) s += cpart[((uint64_t)b * nparts + q) * 32 + cidx];
  cvec[(uint64_t)b * 32 + cidx] = (float)s;
  if (cidx == 0) {
    const float m = __uint_as_float(amax_bits[b]);
    scales[2 * b + 0] = (m > 0.0f) ? 32512.0f / m : 0.0f;
    scales[2 * b + 1] = (m > 0.0f) ? m / 32512.0f : 0.0f;
    amax_bits[b] = 0u;
  }
}

// grid = (chunks, blocks); same image layout as prep_b_i8_kernel, written at the block's image stages
__global__ void __launch_bounds__(256) prep_b_i8_batch_kernel(const float* __restrict__ bin, uint32_t ld,
                                                              const SketchBatchBlock* __restrict__ blocks,
                                                              const float* __restrict__ f,
                                                              const float* __restrict__ scales,
                                                              int8_t* __restrict__ img_all) {
  const SketchBatchBlock bk = blocks[blockIdx.y];
  const float* src = bin + bk.bin_off;
  const float* fb = f ? f + bk.fe_off : nullptr;
  const uint64_t K = bk.K;
  const uint32_t l = bk.l;
  const uint64_t Kpad = (uint64_t)bk.nst * STAGE_FIELDS;
  int8_t* img = img_all + (size_t)bk.img_st0 * B_STAGE_BYTES;
  const float qs = scales[2 * blockIdx.y];
  const uint64_t total = (Kpad / 16) * NL;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t n = (uint32_t)(t % NL);
    const uint64_t kc = t / NL;
    const uint64_t g = kc >> 1;
    const uint32_t half_idx = (uint32_t)(kc & 1);
    __align__(16) int8_t vhi[16], vlo[16];
#pragma unroll
    for (int ss = 0; ss < 16; ++ss) {
      const int s = half_idx * 16 + ss;
      const int cc = s >> 2, bb = s & 3;
      const uint64_t k = g * 32 + 16 * (cc >> 2) + (cc & 3) + 4 * bb;
      int q = 0;
      if (k < K && k >= bk.kskip && n < l) {
        float v = src[k * ld + n];
        if (fb) v *= fb[k];
        q = __float2int_rn(v * qs);
        q = max(-32512, min(32512, q));
      }
      const int h = (q + 128) >> 8;
      vhi[ss] = (int8_t)h;
      vlo[ss] = (int8_t)(q - 256 * h);
    }
    const uint64_t base = g * 32 * NM + (uint64_t)half_idx * (NM * 16);
    *reinterpret_cast<uint4*>(img + base + (n >> 3) * 128 + (n & 7) * 16) = *reinterpret_cast<const uint4*>(vhi);
    *reinterpret_cast<uint4*>(img + base + ((n + 32) >> 3) * 128 + (n & 7) * 16) = *reinterpret_cast<const uint4*>(vlo);
  }
}

CUtensorMapL2promotion l2_promotion() {
  const char* e = getenv("GPCA_DEBUG_L2PROMO");
  const int v = e ? atoi(e) : 0;
  return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                  : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_i8() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
}  // namespace

bool sketch_i8_supported(gpca_ctx* c, const SketchProblem& p) {
  (void)c;
  if (p.l == 0 || p.l > 32) return false;
  if (p.G.rows < 128 || p.G.cols < 256) return false;
  if (p.G.pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(p.G.p) & 15) != 0) return false;
  if (p.G.pitch >= (1ull << 31) || p.G.rows >= (1ull << 31)) return false;
  return get_encode_fn_i8() != nullptr;
}

int launch_sketch_i8(gpca_ctx* c, const SketchProblem& p) {
  const uint64_t K = p.G.cols, rows = p.G.rows;
  const uint64_t Kpad = round_up(K, STAGE_FIELDS);
  const uint32_t total_stages = (uint32_t)(Kpad / STAGE_FIELDS);
  GPCA_CUDA_TRY(c, c->ws_bytes.alloc(Kpad * NM));
  int8_t* img = reinterpret_cast<int8_t*>(c->ws_bytes.p);
  int nb = (int)((K + 31) / 32);
  if (nb > c->sm_count * 8) nb = c->sm_count * 8;
  if (nb < 1) nb = 1;
  GPCA_CUDA_TRY(c, c->ws_cpart.alloc((size_t)nb * 64));
  GPCA_CUDA_TRY(c, c->ws_cvec.alloc(64 + 8));
  float* cvec = c->ws_cvec.p;
  float* scales = c->ws_cvec.p + 64;
  unsigned int* amax = reinterpret_cast<unsigned int*>(c->ws_cvec.p + 66);
  if (!c->tc_amax_zeroed) {
    GPCA_CUDA_TRY(c, cudaMemsetAsync(amax, 0, sizeof(unsigned int), c->stream));
    c->tc_amax_zeroed = true;
  }
  double* st_cpart = nullptr;
  unsigned int* st_amax = nullptr;
  GPCA_TRY(stats_buffer(c, &st_cpart, &st_amax));
  bool have_stats = p.use_stats && c->stats_for == p.Bin && c->stats_l == p.l;
  if (getenv("GPCA_DEBUG_NO_USE_STATS")) have_stats = false;
  if (p.gen) {
    i8_set_scale_kernel<<<1, 1, 0, c->stream>>>(p.gen_amax, scales);
  } else if (have_stats) {
    // the producer of Bin left max |f o Bin| behind (SketchProblem::emit_stats): one sweep over the operand saved
    i8_scale_kernel<<<1, 1, 0, c->stream>>>(st_amax, scales);
    c->stats_pending = false;    // (the scale kernel resets the max-abs word)
  } else {
    i8_amax_kernel<<<nb, 256, 0, c->stream>>>(p.Bin, K, p.l, p.ld, p.f, amax);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
    i8_scale_kernel<<<1, 1, 0, c->stream>>>(amax, scales);
  }
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  c->stats_for = nullptr;
  {
    const uint64_t total = (Kpad / 16) * NL;
    const uint64_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (uint64_t)c->sm_count * 8 ? blocks : (uint64_t)c->sm_count * 8);
    GPCA_CUDA_TRY(c, c->ws_cpart.alloc((size_t)grid * 64));
    if (p.gen)
      prep_b_i8_gauss_kernel<<<grid, 256, 0, c->stream>>>(p.gen_seed, p.gen_stream, p.gen_row0, K, Kpad, p.l, p.f, p.e,
                                                          scales, img, c->ws_cpart.p);
    else
      prep_b_i8_kernel<<<grid, 256, 0, c->stream>>>(p.Bin, K, Kpad, p.l, p.ld, p.f, p.e, scales, img, c->ws_cpart.p);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
    i8_cvec_kernel<<<1, 256, 0, c->stream>>>(c->ws_cpart.p, grid, cvec);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  // K split: the items are dealt round-robin to the persistent CTAs, so a launch lasts ceil(items / slots) item times.
  // Pick the split that minimises rounds x stages per item plus the per-item epilogue and the split-K partial traffic,
  // with at least four items per CTA.  Measured on one box against the old "about four items per CTA" rule: -2.5 % on
  // the snp-side pass of the config-4 shard (6 splits, 2,052 items, 7 full rounds instead of 4 splits / 5 rounds at
  // 92 %); neutral at config 3, where the old rule left six CTAs with a fifth item -- the stragglers run alone on their
  // SMs and at a higher clock, so the tail costs far less than its share of the items.
  // The plan is made for both CTA shapes (RT = 2: 256 rows, two CTAs per SM; RT = 4: 512 rows, one CTA per SM).
  struct Plan {
    uint32_t ksplit, row_groups, slots;
    double est_us;
  };
  auto make_plan = [&](int rt, bool deep_cta = false) {
    Plan pl;
    pl.row_groups = (uint32_t)((rows + rt * 128 - 1) / (rt * 128));
    pl.slots = (uint32_t)c->sm_count * ((rt == 2 && !deep_cta) ? 2u : 1u);
    pl.ksplit = 1;
    const uint32_t max_split = std::max<uint32_t>(1u, (total_stages + 7) / 8);
    const uint32_t hi = std::min<uint32_t>(max_split, (8 * pl.slots + pl.row_groups - 1) / pl.row_groups + 1);
    // one stage of every resident CTA's rows at ~4.2 TB/s
    const double t_stage_us = (double)c->sm_count * (deep_cta ? 1.0 : 2.0) * 16384.0 / 4.2e6;
    const double t_item_us = (rt == 2 && !deep_cta) ? 2.0 : 4.0;   // epilogue / hand-over per item (not hidden behind a second CTA)
    double best = 1e300;
    for (uint32_t ks = 1; ks <= hi; ++ks) {
      const uint32_t spp_c = (total_stages + ks - 1) / ks;
      const uint32_t ks_eff = (total_stages + spp_c - 1) / spp_c;
      if (ks_eff != ks) continue;                                        // same split as a smaller candidate
      const uint64_t items = (uint64_t)pl.row_groups * ks;
      const uint64_t rounds = (items + pl.slots - 1) / pl.slots;
      if (rounds < 4 && ks < hi) continue;      // at least four items per CTA when the K range allows it
      double est = (double)rounds * ((double)spp_c * t_stage_us + t_item_us);
      if (ks > 1) est += (double)ks * (double)rows * 128.0 * 2.0 / 3.0e6;  // partials written and read back (~3 TB/s)
      if (est < best * 0.999) {
        best = est;
        pl.ksplit = ks;
      }
    }
    pl.est_us = best;
    return pl;
  };
  const Plan plan2 = make_plan(2), plan4 = make_plan(4);
  // RT = 4 when its items are long (the epilogue is a small share) and its schedule is not worse by more than the
  // operand traffic it saves (measured: DESIGN.md section 4)
  const uint32_t spp4 = (total_stages + plan4.ksplit - 1) / plan4.ksplit;
  // Measured (profiles/r2_wide_cta_ab.txt, same box, alternating): RT = 4 is SLOWER -- 3.10 vs 2.55 ms at the config-4
  // shard (500,000 x 87,500), 25.0 vs 23.0 ms at 500,000 x 700,000, 1.95 vs 1.63 ms at 2,504 x 10M.  The image stream
  // out of L2 is not what limits the pass; two independent CTA pipelines per SM hide each other's hand-overs, and a
  // hand-over that waits for 16 warps instead of 8 costs more than the halved operand traffic saves.  The shape stays
  // available for experiments (GPCA_I8_WIDE=1) and is kept bit-identical to the regular one by the tests.
  (void)spp4;
  bool wide = false, deep = false;
  if (const char* e = getenv("GPCA_I8_WIDE")) wide = atoi(e) != 0 && rows >= 512;
  if (const char* e = getenv("GPCA_I8_DEEP")) deep = atoi(e) != 0 && !wide;
  const Plan plan2d = make_plan(2, true);
  const Plan& plan = wide ? plan4 : deep ? plan2d : plan2;
  const uint32_t row_groups = plan.row_groups, slots = plan.slots;
  uint32_t ksplit = plan.ksplit;
  if (getenv("GPCA_DEBUG_OLD_KSPLIT")) {      // the rule before the cost model (A/B)
    ksplit = 1;
    if (row_groups < 4 * slots) {
      ksplit = (4 * slots + row_groups - 1) / row_groups;
      const uint32_t max_split = (total_stages + 7) / 8;
      if (ksplit > max_split) ksplit = max_split;
      if (ksplit < 1) ksplit = 1;
    }
  }
  if (const char* dbg = getenv("GPCA_DEBUG_KSPLIT")) ksplit = (uint32_t)atoi(dbg);
  const uint32_t min_split = (total_stages + MAX_STAGES_PER_ITEM - 1) / MAX_STAGES_PER_ITEM;   // int32 headroom
  if (ksplit < min_split) ksplit = min_split;
  const uint32_t spp = (total_stages + ksplit - 1) / ksplit;
  ksplit = (total_stages + spp - 1) / spp;
  const uint64_t n_items64 = (uint64_t)row_groups * ksplit;
  if (n_items64 > 0x7fffffffull) {
    c->set_error("sketch_i8: too many work items");
    return GPCA_ERR_INVALID;
  }
  I8Params tp;
  tp.bimg = img;
  tp.rows = rows;
  tp.total_stages = total_stages;
  tp.stages_per_split = spp;
  tp.ksplit = ksplit;
  tp.row_groups = row_groups;
  tp.n_items = (uint32_t)n_items64;
  tp.a = p.a;
  tp.b = p.b;
  tp.cvec = cvec;
  tp.scales = scales;
  tp.out = p.out;
  tp.ldo = p.ldo;
  tp.l = p.l;
  tp.partial = nullptr;
  tp.items = nullptr;
  const bool emit = p.emit_stats && !c->any_missing && !getenv("GPCA_DEBUG_NO_EMIT_STATS");
  if (emit) GPCA_TRY(stats_begin_produce(c, st_amax));
  tp.stat_amax = (emit && ksplit == 1) ? st_amax : nullptr;
  tp.pad_cols = p.out_pad ? std::min<uint32_t>(p.ldo, (p.l + 7u) & ~7u) : 0;
  if (ksplit > 1) {
    GPCA_CUDA_TRY(c, c->ws_partial.alloc((size_t)ksplit * rows * NL));
    tp.partial = c->ws_partial.p;
  }
  CUtensorMap tmap;
  {
    EncodeTiledFn enc = get_encode_fn_i8();
    const cuuint64_t dims[2] = {(cuuint64_t)(p.G.avail ? p.G.avail : p.G.pitch), (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)p.G.pitch};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)p.G.p, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      c->set_error("sketch_i8: cuTensorMapEncodeTiled failed");
      return GPCA_ERR_CUDA;
    }
  }
  uint32_t grid = tp.n_items < slots ? tp.n_items : slots;
  if (const char* dbg = getenv("GPCA_DEBUG_GRID")) grid = (uint32_t)atoi(dbg);
  KernelTimer kt(c);
  if (wide) {
    constexpr int smem4 = smem_bytes_for<4>();
    const char* tsv = getenv("GPCA_I8_TILE_SYNC");
    if (tsv && atoi(tsv) != 0) {
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 4, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
      sketch_i8_kernel<false, 4, true, 128><<<grid, ACfg<128, 4>::NUM_THREADS, smem4, c->stream>>>(tmap, tp);
    } else {
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 4, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
      sketch_i8_kernel<false, 4, false, 128><<<grid, ACfg<128, 4>::NUM_THREADS, smem4, c->stream>>>(tmap, tp);
    }
  } else if (deep) {
    constexpr int smemd = smem_bytes_for<2, true>();
    const char* tsv = getenv("GPCA_I8_TILE_SYNC");
    if (tsv && atoi(tsv) != 0) {      // + one issuer per row tile
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 2, true, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemd));
      sketch_i8_kernel<false, 2, true, 128, true><<<grid, ACfg<128, 2, true>::NUM_THREADS, smemd, c->stream>>>(tmap, tp);
    } else {
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 2, false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemd));
      sketch_i8_kernel<false, 2, false, 128, true><<<grid, ACfg<128, 2, true>::NUM_THREADS, smemd, c->stream>>>(tmap, tp);
    }
  } else {
    int smem_bytes = smem_bytes_for<2>();
    if (const char* dbg = getenv("GPCA_DEBUG_SMEM_EXTRA")) smem_bytes += atoi(dbg);
    const char* tsv = getenv("GPCA_I8_TILE_SYNC");
    const bool tile_sync = tsv ? atoi(tsv) != 0 : GPCA_I8_TILE_SYNC_DEFAULT;
    if (tile_sync) {
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 2, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      sketch_i8_kernel<false, 2, true, 128><<<grid, ACfg<128, 2>::NUM_THREADS, smem_bytes, c->stream>>>(tmap, tp);
    } else {
      GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<false, 2, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      sketch_i8_kernel<false, 2, false, 128><<<grid, ACfg<128, 2>::NUM_THREADS, smem_bytes, c->stream>>>(tmap, tp);
    }
  }
  kt.end(rows, K, ksplit | (wide ? 0x10000u : 0u) | (deep ? 0x20000u : 0u), tp.n_items);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  if (ksplit > 1) {
    const uint64_t total = rows * NL;
    const uint64_t blocks = (total + 255) / 256;
    const int g2 = (int)(blocks < (uint64_t)c->sm_count * 8 ? blocks : (uint64_t)c->sm_count * 8);
    sketch_reduce_i8_kernel<<<g2, 256, 0, c->stream>>>(tp.partial, (int)ksplit, rows, p.a, p.b, cvec, p.out, p.ldo, p.l,
                                                       emit ? st_amax : nullptr);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  if (emit) {
    c->stats_for = p.out;
    c->stats_l = p.l;
  }
  return GPCA_OK;
}

// ---- batched launch -------------------------------------------------------------------------------------------------
bool sketch_i8_batch_supported(gpca_ctx* c) {
  (void)c;
  return get_encode_fn_i8() != nullptr;
}

int launch_sketch_i8_batch(gpca_ctx* c, const SketchBatch& sb) {
  if (sb.n_items == 0 || sb.n_blocks == 0) return GPCA_OK;
  bool wide_boxes = sb.wide_boxes && (sb.G.avail ? sb.G.avail : sb.G.pitch) >= 128;
  if (const char* e = getenv("GPCA_I8_ITEM_BOX")) wide_boxes = atoi(e) == 128 && (sb.G.avail ? sb.G.avail : sb.G.pitch) >= 128;
  if (sb.G.pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(sb.G.p) & 15) != 0 || sb.G.rows < 128 ||
      (sb.G.avail ? sb.G.avail : sb.G.pitch) < 64) {
    c->set_error("sketch_i8_batch: unsupported matrix shape");
    return GPCA_ERR_INVALID;
  }
  GPCA_CUDA_TRY(c, c->ws_bytes.alloc((size_t)sb.total_img_stages * B_STAGE_BYTES));
  int8_t* img = reinterpret_cast<int8_t*>(c->ws_bytes.p);
  int np = (int)((sb.max_K + 255) / 256);          // 32 rows of 8 per CTA at least
  const int np_cap = (c->sm_count * 16 + (int)sb.n_blocks - 1) / (int)sb.n_blocks;
  if (np > np_cap) np = np_cap;
  if (np < 1) np = 1;
  GPCA_CUDA_TRY(c, c->ws_cpart.alloc((size_t)np * sb.n_blocks * 32));
  // per-block column sums [B][32], scales [B][2], amax bits [B]
  const size_t need = (size_t)sb.n_blocks * (32 + 2 + 1);
  if (c->ws_bstat.n < need) {
    GPCA_CUDA_TRY(c, c->ws_bstat.alloc(need));
    GPCA_CUDA_TRY(c, cudaMemsetAsync(c->ws_bstat.p, 0, need * sizeof(float), c->stream));
  }
  // layout is by the buffer's capacity so that a smaller later batch finds its amax words zeroed
  const size_t cap_blocks = c->ws_bstat.n / 35;
  float* cvec = c->ws_bstat.p;
  float* scales = c->ws_bstat.p + cap_blocks * 32;
  unsigned int* amax = reinterpret_cast<unsigned int*>(c->ws_bstat.p + cap_blocks * 34);
  {
    dim3 g1((unsigned)np, sb.n_blocks);
    i8_colstats_batch_kernel<<<g1, 256, 0, c->stream>>>(sb.Bin, sb.ld, sb.d_blocks, sb.f, sb.e, c->ws_cpart.p, amax);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
    i8_finalize_stats_batch_kernel<<<(sb.n_blocks + 7) / 8, 256, 0, c->stream>>>(c->ws_cpart.p, np, sb.n_blocks, cvec,
                                                                                amax, scales);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
    const uint64_t chunks = ((uint64_t)((sb.max_K + 255) / 256) * 256 / 16 * NL + 255) / 256;
    unsigned gx = (unsigned)(chunks < 64 ? chunks : 64);
    if (gx < 1) gx = 1;
    dim3 g2(gx, sb.n_blocks);
    prep_b_i8_batch_kernel<<<g2, 256, 0, c->stream>>>(sb.Bin, sb.ld, sb.d_blocks, sb.f, scales, img);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  I8Params tp;
  tp.bimg = img;
  tp.rows = sb.G.rows;
  tp.total_stages = sb.total_img_stages;
  tp.stages_per_split = 0;
  tp.ksplit = 1;
  tp.row_groups = 1;
  tp.n_items = sb.n_items;
  tp.a = sb.a;
  tp.b = sb.b;
  tp.cvec = cvec;
  tp.scales = scales;
  tp.out = sb.out;
  tp.ldo = sb.ldo;
  tp.l = 0;
  tp.partial = nullptr;
  tp.items = sb.d_items;
  tp.stat_amax = nullptr;
  tp.pad_cols = 0;
  CUtensorMap tmap;
  {
    EncodeTiledFn enc = get_encode_fn_i8();
    const cuuint64_t dims[2] = {(cuuint64_t)(sb.G.avail ? sb.G.avail : sb.G.pitch), (cuuint64_t)sb.G.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)sb.G.pitch};
    // Box width: items of one or two stages are bound by their own latency and keep 64-byte boxes (four stages in
    // flight); long items on a matrix whose row pitch is tens of KB and more (the condensed-feature pass at 500,000 x
    // 700,000: 7-stage items, 175 KB pitch) are bound by the TMA stream, which 64-byte boxes cap at about half of what
    // 128-byte boxes reach -- one DRAM page per box (tools/probe/tma_bw_probe.cu).
    const cuuint32_t box[2] = {wide_boxes ? 128u : 64u, 128};
    const cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)sb.G.p, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, wide_boxes ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      c->set_error("sketch_i8_batch: cuTensorMapEncodeTiled failed");
      return GPCA_ERR_CUDA;
    }
  }
  constexpr int SMEM_BYTES = smem_bytes_for<2>();
  constexpr int NUM_THREADS = ACfg<64, 2>::NUM_THREADS;
  const uint32_t slots = (uint32_t)c->sm_count * 2;
  const uint32_t grid = tp.n_items < slots ? tp.n_items : slots;
  KernelTimer kt(c);
  if (wide_boxes) {
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<true, 2, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sketch_i8_kernel<true, 2, false, 128><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(tmap, tp);
  } else {
    GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_i8_kernel<true, 2, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sketch_i8_kernel<true, 2, false, 64><<<grid, NUM_THREADS, SMEM_BYTES, c->stream>>>(tmap, tp);
  }
  kt.end();
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  return GPCA_OK;
}
