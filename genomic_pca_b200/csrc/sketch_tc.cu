// sketch_tc.cu -- tcgen05 / TMEM / TMA engine for the sketch pass (engine 1).
//
//   acc[r, :] = sum_k code(r,k) * B'[k, :]          (then out = a_r*scale*acc - b_r*cvec in the epilogue)
//
// Design (DESIGN.md section "K-b"):
//   * the packed 2-bit rows are staged by TMA (cp.async.bulk.tensor, 64-byte boxes, SWIZZLE_64B) and are
//     NEVER expanded through shared memory: each expander thread owns one row (= one TMEM lane), reads 16 B
//     (64 fields) with one conflict-free LDS.128, turns them into 32 registers of fp16x2 with one LOP3 per
//     register, and writes them straight into TMEM (tcgen05.st) as the A operand of tcgen05.mma (A from TMEM).
//   * expansion trick: a 2-bit dosage code c placed anywhere in the mantissa of an fp16 SUBNORMAL is the
//     exact value c * 2^-24 * 4^j (j = field position / 2).  So `w & (0x00030003 << 2j)` IS a pair of fp16
//     operands; the position-dependent factor 4^j is cancelled by pre-scaling the matching row of the dense
//     operand B' by 4^-j (a power of two: exact), and the K order inside one MMA is permuted the same way on
//     both operands (pairs are (field j, field j+8) of a 32-bit word).
//   * B' (the small dense operand, fp16, pre-permuted/pre-scaled into the UMMA canonical K-major core-matrix
//     image by prep_b_tc_kernel) is streamed with plain 1-D bulk copies and read by the MMA from smem.
//   * accumulators (fp32) live in TMEM; 2 row tiles of 128 rows per CTA; 2 CTAs per SM (256 TMEM columns each).
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//     warps 2..9 = expanders (4 per row tile) which also run the epilogue (tcgen05.ld -> scale -> store).
//   * persistent CTAs, static round-robin over (row group, K split) work items.
#include <cuda.h>
#include <cuda_fp16.h>

#include "sketch_tc.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int RT = 2;             // row tiles (of 128 rows) per CTA
constexpr int STAGE_FIELDS = 256; // K fields per pipeline stage (64 B per row)
constexpr int CHUNKS = 4;         // 64-field chunks per stage
constexpr int A_TILE_BYTES = 128 * 64;
constexpr int NUM_THREADS = 384;  // 12 warps: a multiple of 4, so that warp & 3 is the TMEM lane quarter of the warp for every co-resident CTA (warp 11 idles)

template <int NC>
struct Cfg {
  static constexpr int SA = (NC == 32) ? 4 : 3;             // packed-A ring depth (the HBM stream)
  static constexpr int SB = (NC == 32) ? 3 : 2;             // B' ring depth (L2-resident operand)
  static constexpr int SLOTS = (NC == 32) ? 3 : 2;          // TMEM A slots (each holds both row tiles)
  static constexpr int A_STAGE_BYTES = RT * A_TILE_BYTES;
  static constexpr int B_STAGE_BYTES = STAGE_FIELDS * NC * 2;
  static constexpr int TMEM_COLS = 256;
  static constexpr int D_COL0 = 0;                          // accumulators: RT * NC columns
  static constexpr int A_COL0 = RT * NC;                    // A slots: RT * SLOTS * 32 columns
  static_assert(RT * NC + RT * SLOTS * 32 <= TMEM_COLS, "TMEM budget");
  static constexpr int SMEM_BYTES = SA * A_STAGE_BYTES + SB * B_STAGE_BYTES + 256 /*barriers*/ + 320 /*cvec, scale*/;
};

using namespace tcptx;

// 16 fields of a 32-bit word -> 8 registers of fp16x2 subnormals (column c holds fields (c, c+8);
// scale class of column c = {0,1,2,3,4,2,3,4}[c], value = code * 2^-24 * 4^class)
__device__ __forceinline__ void expand_word(uint32_t w, uint32_t* r) {
  const uint32_t u = __umulhi(w, 1u << 26);   // = w >> 6, but on the FMA pipe (IMAD.HI): the ALU pipe is the co-limiter
  r[0] = w & 0x00030003u;
  r[1] = w & 0x000C000Cu;
  r[2] = w & 0x00300030u;
  r[3] = w & 0x00C000C0u;
  r[4] = w & 0x03000300u;
  r[5] = u & 0x00300030u;
  r[6] = u & 0x00C000C0u;
  r[7] = u & 0x03000300u;
}

struct TcParams {
  const __half* bimg;       // [total_stages][STAGE_FIELDS * NC] halfs (UMMA core-matrix image)
  uint64_t rows;
  uint32_t total_stages;
  uint32_t stages_per_split;
  uint32_t ksplit;
  uint32_t row_groups;
  uint32_t n_items;
  const float* a;
  const float* b;
  const float* cvec;
  const float* scales;      // [0] = b_scale (2^beta), [1] = acc_scale (2^(24-beta))
  float* out;
  uint32_t ldo;
  uint32_t l;
  float* partial;           // [ksplit][rows][NC] when ksplit > 1
};

template <int NC>
__global__ void __launch_bounds__(NUM_THREADS, 2)
sketch_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
  using C = Cfg<NC>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  // [A ring: SA x (RT tiles of 128 rows x 64 B)][B ring: SB x B' stage image][barriers]
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = smem_base + C::SA * C::A_STAGE_BYTES;
  const uint32_t bars = b_ring + C::SB * C::B_STAGE_BYTES;
  // barrier map (8 B each)
  auto bar_afull = [&](int s) { return bars + 8u * s; };
  auto bar_aempty = [&](int s) { return bars + 8u * (4 + s); };
  auto bar_bfull = [&](int s) { return bars + 8u * (8 + s); };
  auto bar_bempty = [&](int s) { return bars + 8u * (24 + s); };
  auto bar_tfull = [&](int j) { return bars + 8u * (12 + j); };
  auto bar_tempty = [&](int j) { return bars + 8u * (20 + j); };
  const uint32_t bar_accfull = bars + 8u * 28;
  const uint32_t bar_accempty = bars + 8u * 29;
  const uint32_t tmem_slot = bars + 8u * 30;   // 4 bytes: TMEM base address

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::SA; ++s) {
      mbar_init(bar_afull(s), 1);
      mbar_init(bar_aempty(s), 4 * RT);
    }
    for (int s = 0; s < C::SB; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_bempty(s), 1);
    }
    for (int j = 0; j < C::SLOTS; ++j) {
      mbar_init(bar_tfull(j), 4 * RT);   // every expander warp of both row tiles
      mbar_init(bar_tempty(j), 1);       // tcgen05.commit
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  // epilogue constants staged once per CTA: cvec[64] and the accumulator scale
  const uint32_t cvec_smem = bars + 256;
  if (threadIdx.x >= 64 && threadIdx.x < 128) {
    const int i = threadIdx.x - 64;
    reinterpret_cast<float*>(smem_raw + (cvec_smem - smem_base))[i] = (i < NC) ? p.cvec[i] : 0.0f;
  } else if (threadIdx.x == 128) {
    reinterpret_cast<float*>(smem_raw + (cvec_smem - smem_base))[64] = p.scales[1];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const float s_scale = reinterpret_cast<const float*>(smem_raw + (cvec_smem - smem_base))[64];
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 || warp == 10) {
    // ===================== producers (ONE elected thread runs the whole role: an elected region per stage costs
    // ~100 cycles of divergence / reconvergence each, tools/probe/mma_sttm_probe.cu) ==========
    // warp 0 streams the packed genotype tiles (TMA 2-D, evict-first: read once from HBM);
    // warp 10 streams the B' image (1-D bulk copies, evict-last: shared by all CTAs through L2).
    const bool is_a = (warp == 0);
    uint32_t it = 0;
    const bool leader = elect_one();
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x) {
      const uint32_t ks = item / p.row_groups, rg = item - ks * p.row_groups;
      const uint32_t st0 = ks * p.stages_per_split;
      uint32_t st1 = st0 + p.stages_per_split;
      if (st1 > p.total_stages) st1 = p.total_stages;
      const int row0 = (int)(rg * (RT * 128));
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        if (is_a) {
          const int s = it % C::SA;
          const uint32_t ph = (it / C::SA) & 1u;
          mbar_wait(bar_aempty(s), ph ^ 1u);
          const uint32_t sbase = a_ring + s * C::A_STAGE_BYTES;
          mbar_arrive_expect_tx(bar_afull(s), C::A_STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < RT; ++t)
            tma_load_2d(sbase + t * A_TILE_BYTES, &tmap, bar_afull(s), (int)(st * 64), row0 + t * 128);
        } else {
          const int s = it % C::SB;
          const uint32_t ph = (it / C::SB) & 1u;
          mbar_wait(bar_bempty(s), ph ^ 1u);
          mbar_arrive_expect_tx(bar_bfull(s), C::B_STAGE_BYTES);
          bulk_load_1d(b_ring + s * C::B_STAGE_BYTES, p.bimg + (size_t)st * STAGE_FIELDS * NC, C::B_STAGE_BYTES,
                       bar_bfull(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread runs the whole role) ============
    // instruction descriptor: D=f32, A=B=f16, K-major both, N=NC, M=128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | (8u << 24);
    // B smem descriptor: K-major, no swizzle; LBO = NC*16 B (next 8-wide K chunk), SBO = 128 B (next 8 n), version 1
    const uint32_t desc_lo_const = (uint32_t)((NC * 16) >> 4) << 16;
    const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);
    uint32_t it = 0, cit = 0, item_idx = 0;
    const bool leader = elect_one();
    for (uint32_t item = blockIdx.x; leader && item < p.n_items; item += gridDim.x, ++item_idx) {
      const uint32_t ks = item / p.row_groups;
      const uint32_t st0 = ks * p.stages_per_split;
      uint32_t st1 = st0 + p.stages_per_split;
      if (st1 > p.total_stages) st1 = p.total_stages;
      mbar_wait(bar_accempty, (item_idx & 1u) ^ 1u);   // previous item's epilogue drained the accumulators
      tc_fence_after();
      uint32_t acc_flag = 0;
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const int s = it % C::SB;
        const uint32_t ph = (it / C::SB) & 1u;
        mbar_wait(bar_bfull(s), ph);
        const uint32_t bsm = b_ring + s * C::B_STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < CHUNKS; ++q, ++cit) {
          const int slot = cit % C::SLOTS;
          const uint32_t sph = (cit / C::SLOTS) & 1u;
          mbar_wait(bar_tfull(slot), sph);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < RT; ++t) {
            const uint32_t d_t = tmem_base + C::D_COL0 + t * NC;
            const uint32_t a_t = tmem_base + C::A_COL0 + (slot * RT + t) * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t baddr = bsm + (uint32_t)((q * 4 + i) * (32 * NC));
              const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo_const | ((baddr >> 4) & 0x3FFFu));
              tc_mma_ts(d_t, a_t + 8 * i, bdesc, idesc, acc_flag | (uint32_t)i);
            }
          }
          tc_commit(bar_tempty(slot));
          acc_flag = 1;
        }
        tc_commit(bar_bempty(s));
      }
      tc_commit(bar_accfull);
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 10) {
    // ===================== expanders + epilogue =====================
    const int tile = (warp - 2) >> 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t sw = (uint32_t)((row_in_tile >> 1) & 3);
    uint32_t it = 0, cit = 0, item_idx = 0;
    for (uint32_t item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_idx) {
      const uint32_t ks = item / p.row_groups, rg = item - ks * p.row_groups;
      const uint32_t st0 = ks * p.stages_per_split;
      uint32_t st1 = st0 + p.stages_per_split;
      if (st1 > p.total_stages) st1 = p.total_stages;
      for (uint32_t st = st0; st < st1; ++st, ++it) {
        const int s = it % C::SA;
        const uint32_t ph = (it / C::SA) & 1u;
        mbar_wait(bar_afull(s), ph);
        const uint32_t arow = a_ring + s * C::A_STAGE_BYTES + tile * A_TILE_BYTES + row_in_tile * 64;
        uint4 v[CHUNKS];
#pragma unroll
        for (int q = 0; q < CHUNKS; ++q) {
          const uint32_t addr = arow + (((uint32_t)q ^ sw) << 4);
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w)
                       : "r"(addr));
        }
        // two chunks per completion wait: the second chunk's expansion overlaps the first TMEM store in flight,
        // and one tcgen05.wait::st covers both stores before the two slots are published to the MMA issuer
#pragma unroll
        for (int q = 0; q < CHUNKS; q += 2, cit += 2) {
          const int slot0 = cit % C::SLOTS, slot1 = (cit + 1) % C::SLOTS;
          const uint32_t sph0 = (cit / C::SLOTS) & 1u, sph1 = ((cit + 1) / C::SLOTS) & 1u;
          {
            uint32_t r[32];
            expand_word(v[q].x, r + 0);
            expand_word(v[q].y, r + 8);
            expand_word(v[q].z, r + 16);
            expand_word(v[q].w, r + 24);
            mbar_wait(bar_tempty(slot0), sph0 ^ 1u);   // MMAs that read this TMEM slot have completed
            tc_fence_after();
            tmem_st32(tmem_base + lane_addr + C::A_COL0 + (slot0 * RT + tile) * 32, r);
          }
          {
            uint32_t r[32];
            expand_word(v[q + 1].x, r + 0);
            expand_word(v[q + 1].y, r + 8);
            expand_word(v[q + 1].z, r + 16);
            expand_word(v[q + 1].w, r + 24);
            mbar_wait(bar_tempty(slot1), sph1 ^ 1u);
            tc_fence_after();
            tmem_st32(tmem_base + lane_addr + C::A_COL0 + (slot1 * RT + tile) * 32, r);
          }
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_tfull(slot0));
            mbar_arrive(bar_tfull(slot1));
          }
        }
        // The stage is released only now: every register loaded from it has been consumed by the expansions above, so
        // all of this warp's shared-memory reads have completed before TMA may overwrite the stage (releasing right
        // after issuing the loads raced with loads still in flight when two CTAs share an SM).
        if (lane == 0) mbar_arrive(bar_aempty(s));
      }
      // ---- epilogue of this work item ----
      const uint64_t r = (uint64_t)rg * (RT * 128) + tile * 128 + row_in_tile;
      // per-row epilogue factors are fetched before waiting for the accumulators (latency overlaps the MMA drain)
      float ar = 1.0f, br = 1.0f;
      if (!p.partial && r < p.rows) {
        if (p.a) ar = __ldg(p.a + r);
        if (p.b) br = __ldg(p.b + r);
      }
      mbar_wait(bar_accfull, item_idx & 1u);
      tc_fence_after();
      uint32_t acc[NC];
      {
        uint32_t tmp[32];
        tmem_ld32(tmem_base + lane_addr + C::D_COL0 + tile * NC, tmp);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = tmp[i];
        if (NC == 64) {
          tmem_ld32(tmem_base + lane_addr + C::D_COL0 + tile * NC + 32, tmp);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[(NC == 64 ? 32 : 0) + i] = tmp[i];
        }
      }
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty);
      if (r < p.rows) {
        const float scale = s_scale;
        if (p.partial) {
          float* dst = p.partial + ((uint64_t)ks * p.rows + r) * NC;
#pragma unroll
          for (int cidx = 0; cidx < NC; cidx += 4) {
            float4 o;
            o.x = __uint_as_float(acc[cidx + 0]) * scale;
            o.y = __uint_as_float(acc[cidx + 1]) * scale;
            o.z = __uint_as_float(acc[cidx + 2]) * scale;
            o.w = __uint_as_float(acc[cidx + 3]) * scale;
            *reinterpret_cast<float4*>(dst + cidx) = o;
          }
        } else {
          const float as = ar * scale;
          float* dst = p.out + r * p.ldo;
          const float* cv = reinterpret_cast<const float*>(smem_raw + (cvec_smem - smem_base));
          if ((p.ldo & 1u) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0) {
#pragma unroll
            for (int cidx = 0; cidx < NC; cidx += 2) {
              if ((uint32_t)cidx + 1 < p.l) {
                float2 o;
                o.x = as * __uint_as_float(acc[cidx]) - br * cv[cidx];
                o.y = as * __uint_as_float(acc[cidx + 1]) - br * cv[cidx + 1];
                *reinterpret_cast<float2*>(dst + cidx) = o;
              } else if ((uint32_t)cidx < p.l) {
                dst[cidx] = as * __uint_as_float(acc[cidx]) - br * cv[cidx];
              }
            }
          } else {
#pragma unroll
            for (int cidx = 0; cidx < NC; ++cidx)
              if ((uint32_t)cidx < p.l) dst[cidx] = as * __uint_as_float(acc[cidx]) - br * cv[cidx];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- operand preparation --------------------------------------------------------------------
// column statistics of the dense operand: cpart[block][c] = sum_k e_k Bin[k][c] (f64), amax = max |f_k Bin[k][c]|
__global__ void __launch_bounds__(256) tc_colstats_kernel(const float* __restrict__ bin, uint64_t K, uint32_t l,
                                                          uint32_t ld, const float* __restrict__ f,
                                                          const float* __restrict__ e, double* __restrict__ cpart,
                                                          unsigned int* __restrict__ amax_bits) {
  // lane = column (and column + 32 when l > 32), 8 row lanes per CTA, 4 independent rows in flight per thread
  __shared__ double red[2][256];
  __shared__ float redm[256];
  const int cidx = threadIdx.x & 31;
  const int rr = threadIdx.x >> 5;
  const bool c0 = (uint32_t)cidx < l, c1 = (uint32_t)(cidx + 32) < l;
  double acc0 = 0.0, acc1 = 0.0;
  float mx = 0.0f;
  const uint64_t stride = (uint64_t)gridDim.x * 8;
  for (uint64_t k = (uint64_t)blockIdx.x * 8 + rr; k < K; k += 4 * stride) {
    float x0[4], x1[4], ek[4], fk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t kk = k + u * stride;
      const bool live = kk < K;
      x0[u] = (live && c0) ? bin[kk * ld + cidx] : 0.0f;
      x1[u] = (live && c1) ? bin[kk * ld + cidx + 32] : 0.0f;
      ek[u] = (live && e) ? e[kk] : 1.0f;
      fk[u] = (live && f) ? f[kk] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc0 += (double)(x0[u] * ek[u]);
      acc1 += (double)(x1[u] * ek[u]);
      mx = fmaxf(mx, fmaxf(fabsf(x0[u] * fk[u]), fabsf(x1[u] * fk[u])));
    }
  }
  red[0][threadIdx.x] = acc0;
  red[1][threadIdx.x] = acc1;
  redm[threadIdx.x] = mx;
  __syncthreads();
  if (rr == 0) {
    double s0 = 0.0, s1 = 0.0;
    float m = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      s0 += red[0][q * 32 + cidx];
      s1 += red[1][q * 32 + cidx];
      m = fmaxf(m, redm[q * 32 + cidx]);
    }
    cpart[(uint64_t)blockIdx.x * 64 + cidx] = s0;
    cpart[(uint64_t)blockIdx.x * 64 + 32 + cidx] = s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (cidx == 0 && m > 0.0f && isfinite(m)) atomicMax(amax_bits, __float_as_uint(m));
  }
}

__global__ void tc_finalize_stats_kernel(const double* __restrict__ cpart, int nparts, float* __restrict__ cvec,
                                         unsigned int* __restrict__ amax_bits, float* __restrict__ scales) {
  const int cidx = threadIdx.x;
  if (cidx < 64) {
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += cpart[(uint64_t)q * 64 + cidx];
    cvec[cidx] = (float)s;
  }
  if (cidx == 0) {
    const float m = __uint_as_float(*amax_bits);
    int beta = 0;
    if (m > 0.0f) {
      int ex;
      frexpf(m, &ex);          // m = fr * 2^ex, fr in [0.5, 1)  ->  m * 2^(15-ex) < 2^15
      beta = 15 - ex;
    }
    scales[0] = exp2f((float)beta);
    scales[1] = exp2f((float)(24 - beta));
    *amax_bits = 0u;           // ready for the next pass
  }
}

// B' image: for MMA group g (16 consecutive k), K-slot s = 2c + h holds source row 16 g + c + 8 h scaled by
// f_k * 2^beta * 4^-class(c); element (slot s, column n) lives at half offset
//   g*16*NC + (s/8)*(8*NC) + (n/8)*64 + (n%8)*8 + (s%8)          (UMMA K-major core matrices, no swizzle)
template <int NC>
__global__ void __launch_bounds__(256) prep_b_tc_kernel(const float* __restrict__ bin, uint64_t K, uint64_t Kpad,
                                                        uint32_t l, uint32_t ld, const float* __restrict__ f,
                                                        const float* __restrict__ scales, __half* __restrict__ img) {
  const uint64_t total = (Kpad / 8) * NC;   // one thread per (8-slot K chunk, column)
  const float bs = scales[0];
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t n = (uint32_t)(t % NC);
    const uint64_t kc = t / NC;             // global 8-slot chunk index
    const uint64_t g = kc >> 1;
    const uint32_t half_idx = (uint32_t)(kc & 1);
    __align__(16) __half vals[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int s = half_idx * 8 + kk;
      const int cc = s >> 1, h = s & 1;
      const int cls = (cc < 5) ? cc : cc - 3;
      const uint64_t k = g * 16 + cc + 8 * h;
      float v = 0.0f;
      if (k < K && n < l) {
        v = bin[k * ld + n];
        if (f) v *= f[k];
        v *= bs * (1.0f / (float)(1 << (2 * cls)));
      }
      vals[kk] = __float2half_rn(v);
    }
    const uint64_t off = g * 16 * NC + (uint64_t)half_idx * (8 * NC) + (n >> 3) * 64 + (n & 7) * 8;
    *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(vals);
  }
}

template <int NC>
__global__ void sketch_reduce_tc_kernel(const float* __restrict__ partial, int nsplit, uint64_t rows,
                                        const float* __restrict__ a, const float* __restrict__ b,
                                        const float* __restrict__ cvec, float* __restrict__ out, uint32_t ldo,
                                        uint32_t l) {
  const uint64_t total = rows * NC;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / NC;
    const uint32_t cc = (uint32_t)(t % NC);
    if (cc >= l) continue;
    float s = 0.0f;
    for (int q = 0; q < nsplit; ++q) s += partial[((uint64_t)q * rows + r) * NC + cc];
    const float ar = a ? a[r] : 1.0f, br = b ? b[r] : 1.0f;
    out[r * ldo + cc] = ar * s - br * cvec[cc];
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

bool make_tmap(const PackedMat& g, CUtensorMap* map) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  // a K-range view starts inside a physical row: only `avail` bytes belong to it (the rest is zero-filled by TMA)
  const cuuint64_t dims[2] = {(cuuint64_t)(g.avail ? g.avail : g.pitch), (cuuint64_t)g.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)g.pitch};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)g.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NC>
int run_tc(gpca_ctx* c, const SketchProblem& p) {
  using C = Cfg<NC>;
  const uint64_t K = p.G.cols, rows = p.G.rows;
  const uint64_t Kpad = round_up(K, STAGE_FIELDS);
  const uint32_t total_stages = (uint32_t)(Kpad / STAGE_FIELDS);
  // ---- operand prep
  GPCA_CUDA_TRY(c, c->ws_bytes.alloc(Kpad * NC * sizeof(__half)));
  __half* img = reinterpret_cast<__half*>(c->ws_bytes.p);
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
  tc_colstats_kernel<<<nb, 256, 0, c->stream>>>(p.Bin, K, p.l, p.ld, p.f, p.e, c->ws_cpart.p, amax);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  tc_finalize_stats_kernel<<<1, 64, 0, c->stream>>>(c->ws_cpart.p, nb, cvec, amax, scales);
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  {
    const uint64_t total = (Kpad / 8) * NC;
    const uint64_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (uint64_t)c->sm_count * 8 ? blocks : (uint64_t)c->sm_count * 8);
    prep_b_tc_kernel<NC><<<grid, 256, 0, c->stream>>>(p.Bin, K, Kpad, p.l, p.ld, p.f, scales, img);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  // ---- work decomposition
  const uint32_t row_groups = (uint32_t)((rows + RT * 128 - 1) / (RT * 128));
  const uint32_t slots = (uint32_t)c->sm_count * 2;
  uint32_t ksplit = 1;
  if (row_groups < 4 * slots) {
    ksplit = (4 * slots + row_groups - 1) / row_groups;
    const uint32_t max_split = (total_stages + 7) / 8;     // at least 8 stages (2048 fields) per split
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
  }
  const uint32_t spp = (total_stages + ksplit - 1) / ksplit;
  ksplit = (total_stages + spp - 1) / spp;
  const uint64_t n_items64 = (uint64_t)row_groups * ksplit;
  if (n_items64 > 0x7fffffffull) {
    c->set_error("sketch_tc: too many work items");
    return GPCA_ERR_INVALID;
  }
  TcParams tp;
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
  if (ksplit > 1) {
    GPCA_CUDA_TRY(c, c->ws_partial.alloc((size_t)ksplit * rows * NC));
    tp.partial = c->ws_partial.p;
  }
  CUtensorMap tmap;
  if (!make_tmap(p.G, &tmap)) {
    c->set_error("sketch_tc: cuTensorMapEncodeTiled failed");
    return GPCA_ERR_CUDA;
  }
  GPCA_CUDA_TRY(c, cudaFuncSetAttribute(sketch_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const uint32_t grid = tp.n_items < slots ? tp.n_items : slots;
  KernelTimer kt(c);
  sketch_tc_kernel<NC><<<grid, NUM_THREADS, C::SMEM_BYTES, c->stream>>>(tmap, tp);
  kt.end();
  c->launches++;
  GPCA_CUDA_TRY(c, cudaGetLastError());
  if (ksplit > 1) {
    const uint64_t total = rows * NC;
    const uint64_t blocks = (total + 255) / 256;
    const int g2 = (int)(blocks < (uint64_t)c->sm_count * 8 ? blocks : (uint64_t)c->sm_count * 8);
    sketch_reduce_tc_kernel<NC><<<g2, 256, 0, c->stream>>>(tp.partial, (int)ksplit, rows, p.a, p.b, cvec, p.out, p.ldo, p.l);
    c->launches++;
    GPCA_CUDA_TRY(c, cudaGetLastError());
  }
  return GPCA_OK;
}

}  // namespace

bool sketch_tc_supported(gpca_ctx* c, const SketchProblem& p) {
  (void)c;
  if (p.l == 0 || p.l > 64) return false;
  if (p.G.rows < 128 || p.G.cols < 256) return false;      // tiny problems: SIMT engine
  if (p.G.pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(p.G.p) & 15) != 0) return false;   // TMA: 16 B
  if (p.G.pitch >= (1ull << 31) || p.G.rows >= (1ull << 31)) return false;
  return get_encode_fn() != nullptr;
}

int launch_sketch_tc(gpca_ctx* c, const SketchProblem& p) {
  if (p.l <= 32) return run_tc<32>(c, p);
  return run_tc<64>(c, p);
}
