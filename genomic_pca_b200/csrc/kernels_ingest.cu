// kernels_ingest.cu -- PLINK 2-bit ingest: re-pitch/gather, allele-count kernel (K-a),
// recode to the resident dosage coding, 2-bit transpose, standardized-block accessor.
//
// Reference semantics replaced here:
//   * bed-reader reads with count_a1 (src/prepare.rs:622-629, 682-687): 00->2, 01->missing, 10->1, 11->0
//   * pass 1 of perform_snp_qc_and_calc_std_params (src/prepare.rs:1232-1279): integer counts
//   * get_standardized_snp_sample_block (src/prepare.rs:1884-2016)
#include <algorithm>

#include "kernels.cuh"
#include "philox.cuh"

#define KLAUNCH_CHECK(c)                                   \
  do {                                                     \
    (c)->launches++;                                       \
    GPCA_CUDA_TRY((c), cudaGetLastError());                \
  } while (0)

// ------------------------------------------------------------------------------------------
// Re-pitch (and optionally gather samples).  One thread per output 32-bit word (16 fields).
// Source rows may start at any byte offset (pitch ceil(N/4) is rarely 4-aligned), so the source
// is read with byte loads; they hit the same L1 sectors across the warp.
__global__ void repitch_gather_kernel(const uint8_t* __restrict__ in, size_t in_pitch, uint64_t n_in,
                                      const int64_t* __restrict__ keep, uint64_t N, uint64_t M,
                                      uint8_t* __restrict__ out, size_t out_pitch) {
  const uint64_t words_per_row = out_pitch / 4;
  const uint64_t total = M * words_per_row;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = t / words_per_row;
    const uint64_t wi = t - row * words_per_row;
    const uint8_t* src = in + row * in_pitch;
    const uint64_t k0 = wi * 16;
    uint32_t w = 0x55555555u;  // all fields = 01 (missing) -> pads
    if (k0 < N) {
      if (keep == nullptr) {
        uint32_t v = 0;
        const uint64_t b0 = wi * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t byte = 0x55u;
          if ((b0 + j) * 4 < n_in) byte = src[b0 + j];
          v |= byte << (8 * j);
        }
        w = v;
        if (k0 + 16 > N) {  // mask the tail fields to 01
          const int valid = (int)(N - k0);
          const uint32_t m = (valid >= 16) ? 0xffffffffu : ((1u << (2 * valid)) - 1u);
          w = (v & m) | (0x55555555u & ~m);
        }
      } else {
        uint32_t v = 0;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
          uint32_t code = 1u;
          if (k0 + j < N) {
            const uint64_t s = (uint64_t)keep[k0 + j];
            code = (src[s >> 2] >> (2 * (s & 3))) & 3u;
          }
          v |= code << (2 * j);
        }
        w = v;
      }
    }
    *reinterpret_cast<uint32_t*>(out + row * out_pitch + wi * 4) = w;
  }
}

int launch_repitch_gather(gpca_ctx* c, const uint8_t* d_in, size_t in_pitch, uint64_t n_in_samples,
                          const int64_t* d_keep, uint64_t N, uint64_t M, uint8_t* d_out, size_t out_pitch) {
  if (M == 0) return GPCA_OK;
  const uint64_t total = M * (out_pitch / 4);
  const int threads = 256;
  const uint64_t blocks = (total + threads - 1) / threads;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 32 ? blocks : (uint64_t)c->sm_count * 32);
  repitch_gather_kernel<<<grid, threads, 0, c->stream>>>(d_in, in_pitch, n_in_samples, d_keep, N, M, d_out, out_pitch);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// VCF path: variant-major u8 dosages (vcf.rs:293-315 Vec<Vec<u8>>) -> PLINK codes.
__global__ void u8_to_plink_kernel(const uint8_t* __restrict__ in, uint64_t N, uint64_t M,
                                   uint8_t* __restrict__ out, size_t out_pitch) {
  const uint64_t words_per_row = out_pitch / 4;
  const uint64_t total = M * words_per_row;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = t / words_per_row;
    const uint64_t wi = t - row * words_per_row;
    const uint8_t* src = in + row * N;
    const uint64_t k0 = wi * 16;
    uint32_t w = 0;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      uint32_t code = 1u;
      if (k0 + j < N) {
        const uint32_t d = src[k0 + j];
        code = (d == 0) ? 3u : (d == 1) ? 2u : (d == 2) ? 0u : 1u;
      }
      w |= code << (2 * j);
    }
    *reinterpret_cast<uint32_t*>(out + row * out_pitch + wi * 4) = w;
  }
}

int launch_u8_to_plink(gpca_ctx* c, const uint8_t* d_in, uint64_t N, uint64_t M, uint8_t* d_out, size_t out_pitch) {
  if (M == 0) return GPCA_OK;
  const uint64_t total = M * (out_pitch / 4);
  const int threads = 256;
  const uint64_t blocks = (total + threads - 1) / threads;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 32 ? blocks : (uint64_t)c->sm_count * 32);
  u8_to_plink_kernel<<<grid, threads, 0, c->stream>>>(d_in, N, M, d_out, out_pitch);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// K-a  bed_counts: HBM-bound.  128-bit loads, 2-bit field popcounts, shuffle reduction.
//   lo = w & 0x5555.., hi = (w>>1) & 0x5555..;  n(11)=popc(hi&lo)  n(10)=popc(hi&~lo)  n(01)=popc(~hi&lo)
// GROUP lanes cooperate on one row (GROUP = 8, 32) or a whole CTA does (GROUP = 0).
__device__ __forceinline__ void count_word(uint32_t w, uint32_t& c01, uint32_t& c10, uint32_t& c11) {
  const uint32_t lo = w & 0x55555555u;
  const uint32_t hi = (w >> 1) & 0x55555555u;
  c11 += __popc(hi & lo);
  c10 += __popc(hi & ~lo);
  c01 += __popc(lo & ~hi);
}
// Two words per population count.  The class masks of a word have their bits on the even positions only, so the masks
// of a second word fit on the odd positions of the same register: x is classified on the even bits (x & (x >> 1) ...),
// y on the odd bits (y & (y << 1) ...), and one POPC counts both.  POPC issues at a quarter of the ALU rate (16 per
// clock per SM): with one POPC per class and word the kernel sat at 4.5 TB/s, 0.69 of the HBM roofline, POPC-bound;
// this halves the POPCs (six per 16 bytes) for two more logic operations per pair.
__device__ __forceinline__ void count_pair(uint32_t x, uint32_t y, uint32_t& c01, uint32_t& c10, uint32_t& c11) {
  const uint32_t xs = x >> 1, ys = y << 1;
  // even bits: hi = xs, lo = x;  odd bits: hi = y, lo = ys
  const uint32_t m11 = (x & xs & 0x55555555u) | (y & ys & 0xAAAAAAAAu);
  const uint32_t m10 = (xs & ~x & 0x55555555u) | (y & ~ys & 0xAAAAAAAAu);
  const uint32_t m01 = (x & ~xs & 0x55555555u) | (ys & ~y & 0xAAAAAAAAu);
  c11 += __popc(m11);
  c10 += __popc(m10);
  c01 += __popc(m01);
}
__device__ __forceinline__ void count_chunk(const uint4& v, uint32_t& c01, uint32_t& c10, uint32_t& c11) {
  count_pair(v.x, v.y, c01, c10, c11);
  count_pair(v.z, v.w, c01, c10, c11);
}

template <int GROUP>
__global__ void __launch_bounds__(256) bed_counts_kernel(const uint8_t* __restrict__ raw, size_t pitch, uint64_t M,
                                                         uint4* __restrict__ out) {
  const int chunks = (int)(pitch / 16);
  if constexpr (GROUP > 0) {
    // warp-uniform outer loop (all 32 lanes reach every shuffle); a warp covers 32/GROUP rows per step
    constexpr int GPW = 32 / GROUP;
    const uint64_t warp_id = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x % GROUP;
    const int sub = (threadIdx.x & 31) / GROUP;
    for (uint64_t base = warp_id * GPW; base < M; base += nwarps * GPW) {
      const uint64_t row = base + sub;
      const bool live = row < M;
      const uint4* p = reinterpret_cast<const uint4*>(raw + (live ? row : 0) * pitch);
      uint32_t c01 = 0, c10 = 0, c11 = 0;
      for (int i = lane; live && i < chunks; i += GROUP) {
        const uint4 v = ldg_nc_v4(p + i);
        count_chunk(v, c01, c10, c11);
      }
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) {
        c01 += __shfl_xor_sync(0xffffffffu, c01, o);
        c10 += __shfl_xor_sync(0xffffffffu, c10, o);
        c11 += __shfl_xor_sync(0xffffffffu, c11, o);
      }
      if (lane == 0 && live) out[row] = make_uint4(c01, c10, c11, 0u);
    }
  } else {
    __shared__ uint32_t s[3][8];
    for (uint64_t row = blockIdx.x; row < M; row += gridDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(raw + row * pitch);
      uint32_t c01 = 0, c10 = 0, c11 = 0;
      for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
        const uint4 v = ldg_nc_v4(p + i);
        count_chunk(v, c01, c10, c11);
      }
      c01 = warp_sum_u32(c01);
      c10 = warp_sum_u32(c10);
      c11 = warp_sum_u32(c11);
      const int wid = threadIdx.x >> 5;
      if ((threadIdx.x & 31) == 0) {
        s[0][wid] = c01;
        s[1][wid] = c10;
        s[2][wid] = c11;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
          t0 += s[0][w];
          t1 += s[1][w];
          t2 += s[2][w];
        }
        out[row] = make_uint4(t0, t1, t2, 0u);
      }
      __syncthreads();
    }
  }
}

int launch_bed_counts(gpca_ctx* c, const uint8_t* d_raw, size_t pitch, uint64_t M, uint4* d_out) {
  if (M == 0) return GPCA_OK;
  const int threads = 256;
  const int chunks = (int)(pitch / 16);
  // grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
  const int grid_full = c->sm_count * 8;
  if (chunks <= 128) {
    // rows up to 2 KB (8,192 samples): 8 lanes per row, 4 rows per warp -- every lane stays busy and a row costs 9
    // shuffles instead of 15 (at 626-byte rows the warp-per-row variant ran at 1.7 TB/s)
    uint64_t need = (M * 8 + threads - 1) / threads;
    int grid = (int)(need < (uint64_t)grid_full ? need : (uint64_t)grid_full);
    bed_counts_kernel<8><<<grid, threads, 0, c->stream>>>(d_raw, pitch, M, d_out);
  } else if (chunks <= 1024) {
    uint64_t need = (M * 32 + threads - 1) / threads;
    int grid = (int)(need < (uint64_t)grid_full ? need : (uint64_t)grid_full);
    bed_counts_kernel<32><<<grid, threads, 0, c->stream>>>(d_raw, pitch, M, d_out);
  } else {
    int grid = (int)(M < (uint64_t)grid_full ? M : (uint64_t)grid_full);
    bed_counts_kernel<0><<<grid, threads, 0, c->stream>>>(d_raw, pitch, M, d_out);
  }
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Gather PCA rows + recode PLINK -> dosage code:  d_hi = ~hi, d_lo = lo ^ hi
//   00->10 (2)  01->11 (3 = missing)  10->01 (1)  11->00 (0);   pad fields (k >= N) -> 00.
__device__ __forceinline__ uint32_t recode_word(uint32_t w) {
  return ((~w) & 0xAAAAAAAAu) | ((w ^ (w >> 1)) & 0x55555555u);
}

__global__ void build_gs_kernel(const uint8_t* __restrict__ raw, size_t raw_pitch, const uint64_t* __restrict__ idx,
                                uint8_t* __restrict__ gs, size_t gs_pitch, uint64_t D, uint64_t N) {
  const uint64_t chunks_per_row = gs_pitch / 16;
  const uint64_t raw_chunks = raw_pitch / 16;
  const uint64_t total = D * chunks_per_row;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = t / chunks_per_row;
    const uint64_t ci = t - row * chunks_per_row;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ci < raw_chunks && ci * 64 < N) {
      const uint64_t src_row = idx ? idx[row] : row;
      v = ldg_nc_v4(raw + src_row * raw_pitch + ci * 16);
      uint32_t w[4] = {recode_word(v.x), recode_word(v.y), recode_word(v.z), recode_word(v.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t k0 = ci * 64 + (uint64_t)j * 16;
        if (k0 >= N) w[j] = 0;
        else if (k0 + 16 > N) w[j] &= (1u << (2 * (int)(N - k0))) - 1u;
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(gs + row * gs_pitch + ci * 16) = v;
  }
}

int launch_build_gs(gpca_ctx* c, const uint8_t* d_raw, size_t raw_pitch, const uint64_t* d_idx, PackedMat gs) {
  if (gs.rows == 0) return GPCA_OK;
  const uint64_t total = gs.rows * (gs.pitch / 16);
  const int threads = 256;
  const uint64_t blocks = (total + threads - 1) / threads;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  build_gs_kernel<<<grid, threads, 0, c->stream>>>(d_raw, raw_pitch, d_idx, gs.p, gs.pitch, gs.rows, gs.cols);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Streaming ingest (gpca_ingest_bed): the chunk that has just crossed PCIe is counted and recoded straight from its
// staging buffer -- rows at the FILE's pitch ceil(N/4), which is rarely 16-byte aligned -- so no re-pitched copy of
// the payload exists any more (it was a third full-size copy of the matrix next to Gs and Gt, and two more sweeps).
// 16 bytes [16 ci, 16 ci + 16) of a row that starts at any byte address: two aligned 128-bit loads and a funnel
// shift; 2-bit fields at or past `n_fields` read as 00.  The buffer must be readable 16 bytes past its last row.
__device__ __forceinline__ uint4 load_row_chunk(const uint8_t* row, uint64_t n_fields, uint64_t ci) {
  const uintptr_t ua = reinterpret_cast<uintptr_t>(row) + ci * 16;
  const uint4* a0 = reinterpret_cast<const uint4*>(ua & ~(uintptr_t)15);
  const uint32_t sh = (uint32_t)(ua & 15);
  uint4 r = ldg_nc_v4(a0);
  if (sh) {
    const uint4 b = ldg_nc_v4(a0 + 1);
    const uint32_t w[8] = {r.x, r.y, r.z, r.w, b.x, b.y, b.z, b.w};
    const uint32_t ws = sh >> 2, bs = (sh & 3u) * 8u;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = (ws == 0) ? w[i] : (ws == 1) ? w[i + 1] : (ws == 2) ? w[i + 2] : w[i + 3];
    r.x = __funnelshift_r(v[0], v[1], bs);
    r.y = __funnelshift_r(v[1], v[2], bs);
    r.z = __funnelshift_r(v[2], v[3], bs);
    r.w = __funnelshift_r(v[3], v[4], bs);
  }
  const uint64_t f0 = ci * 64;
  if (f0 + 64 > n_fields) {
    uint32_t* rw = &r.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t k0 = f0 + 16u * j;
      if (k0 >= n_fields) rw[j] = 0u;
      else if (k0 + 16 > n_fields) rw[j] &= (1u << (2 * (int)(n_fields - k0))) - 1u;
    }
  }
  return r;
}

// K-a on a staged chunk: out[row] = {n(01), n(10), n(11), 0} over the row's first N fields.  `out` may be mapped
// pinned host memory (the 16-byte records then land on the host without a copy in the transfer queue).
// Counting does not care where a field sits, so the rows -- which start at any byte address -- are walked in ALIGNED
// 16-byte chunks, each loaded exactly once; only the first and the last chunk of a row are masked (bytes of the
// neighbouring rows, fields at or past N).  (The recode, which has to produce aligned rows, needs the two-load funnel
// shift of load_row_chunk; the count kernel with it ran at 3.5 TB/s on 626-byte rows.)
struct RowSpan {
  const uint4* base;     // aligned chunk that holds the row's first byte
  int64_t lo_bit;        // first valid bit, relative to base
  int64_t hi_bit;        // one past the last valid bit (2 N fields further)
  int chunks;            // aligned chunks the row touches
};
__device__ __forceinline__ RowSpan row_span(const uint8_t* row, uint64_t N) {
  RowSpan r;
  const uintptr_t ua = reinterpret_cast<uintptr_t>(row);
  r.base = reinterpret_cast<const uint4*>(ua & ~(uintptr_t)15);
  r.lo_bit = (int64_t)(ua & 15) * 8;
  r.hi_bit = r.lo_bit + 2 * (int64_t)N;
  r.chunks = (int)((r.hi_bit + 127) >> 7);
  return r;
}
__device__ __forceinline__ uint4 span_chunk(const RowSpan& r, int ci) {
  uint4 v = ldg_nc_v4(r.base + ci);
  if (ci == 0 || ci == r.chunks - 1) {
    const int64_t lo = r.lo_bit - 128 * (int64_t)ci, hi = r.hi_bit - 128 * (int64_t)ci;
    uint32_t* w = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t l = lo - 32 * j, h = hi - 32 * j;
      uint32_t m = 0xffffffffu;
      if (h <= 0) m = 0u;
      else if (h < 32) m = (1u << (int)h) - 1u;
      if (l >= 32) m = 0u;
      else if (l > 0) m &= ~((1u << (int)l) - 1u);
      w[j] &= m;
    }
  }
  return v;
}

template <int GROUP>
__global__ void __launch_bounds__(256) chunk_counts_kernel(const uint8_t* __restrict__ src, size_t pitch, uint64_t N,
                                                           uint64_t M, uint4* __restrict__ out) {
  if constexpr (GROUP > 0) {
    constexpr int GPW = 32 / GROUP;
    const uint64_t warp_id = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x % GROUP;
    const int sub = (threadIdx.x & 31) / GROUP;
    for (uint64_t base = warp_id * GPW; base < M; base += nwarps * GPW) {
      const uint64_t row = base + sub;
      const bool live = row < M;
      const RowSpan sp = row_span(src + (live ? row : 0) * pitch, N);
      uint32_t c01 = 0, c10 = 0, c11 = 0;
      for (int i = lane; live && i < sp.chunks; i += GROUP) {
        const uint4 v = span_chunk(sp, i);
        count_chunk(v, c01, c10, c11);
      }
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) {
        c01 += __shfl_xor_sync(0xffffffffu, c01, o);
        c10 += __shfl_xor_sync(0xffffffffu, c10, o);
        c11 += __shfl_xor_sync(0xffffffffu, c11, o);
      }
      if (lane == 0 && live) out[row] = make_uint4(c01, c10, c11, 0u);
    }
  } else {
    __shared__ uint32_t s[3][8];
    for (uint64_t row = blockIdx.x; row < M; row += gridDim.x) {
      const RowSpan sp = row_span(src + row * pitch, N);
      uint32_t c01 = 0, c10 = 0, c11 = 0;
      for (int i = threadIdx.x; i < sp.chunks; i += blockDim.x) {
        const uint4 v = span_chunk(sp, i);
        count_chunk(v, c01, c10, c11);
      }
      c01 = warp_sum_u32(c01);
      c10 = warp_sum_u32(c10);
      c11 = warp_sum_u32(c11);
      const int wid = threadIdx.x >> 5;
      if ((threadIdx.x & 31) == 0) {
        s[0][wid] = c01;
        s[1][wid] = c10;
        s[2][wid] = c11;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
          t0 += s[0][w];
          t1 += s[1][w];
          t2 += s[2][w];
        }
        out[row] = make_uint4(t0, t1, t2, 0u);
      }
      __syncthreads();
    }
  }
}

template <int GROUP>
static void launch_chunk_counts_g(gpca_ctx* c, const uint8_t* d_src, size_t pitch, uint64_t N, uint64_t M, uint4* out) {
  const int threads = 256;
  const int grid_full = c->sm_count * 8;      // a multiple of the SM count, 8 resident CTAs of 256 threads per SM
  if (GROUP == 0) {
    const int grid = (int)(M < (uint64_t)grid_full ? M : (uint64_t)grid_full);
    chunk_counts_kernel<0><<<grid, threads, 0, c->stream>>>(d_src, pitch, N, M, out);
  } else {
    const uint64_t need = (M * (GROUP ? GROUP : 1) + threads - 1) / threads;
    const int grid = (int)(need < (uint64_t)grid_full ? need : (uint64_t)grid_full);
    chunk_counts_kernel<GROUP><<<grid, threads, 0, c->stream>>>(d_src, pitch, N, M, out);
  }
}

int launch_chunk_counts(gpca_ctx* c, const uint8_t* d_src, size_t pitch, uint64_t N, uint64_t M, uint4* out) {
  if (M == 0) return GPCA_OK;
  const int chunks = (int)((N + 63) / 64) + 1;      // aligned 16-byte chunks a row may touch
  // Lanes per row, from a sweep on B200 (profiles/r2_ka_counts_probe.txt): a lane per row for rows of a few chunks
  // (64 samples = one chunk per SNP); otherwise at least four lanes (whole 64-byte runs per row and step) and about 128
  // chunks per lane, up to a warp per row -- which beats a CTA per row even at 125 KB rows (6.26 vs 6.04 TB/s) as long
  // as there are rows for every warp.
  int group = 1;
  if (chunks >= 8) {
    group = 4;
    while (group < 32 && group * 128 < chunks) group *= 2;
  }
  if (chunks > 1024 && M < (uint64_t)c->sm_count * 16) group = 0;      // few, very long rows: a CTA per row
  if (const char* e = getenv("GPCA_DEBUG_COUNT_GROUP")) group = atoi(e);
  switch (group) {
    case 0: launch_chunk_counts_g<0>(c, d_src, pitch, N, M, out); break;
    case 1: launch_chunk_counts_g<1>(c, d_src, pitch, N, M, out); break;
    case 2: launch_chunk_counts_g<2>(c, d_src, pitch, N, M, out); break;
    case 4: launch_chunk_counts_g<4>(c, d_src, pitch, N, M, out); break;
    case 8: launch_chunk_counts_g<8>(c, d_src, pitch, N, M, out); break;
    case 16: launch_chunk_counts_g<16>(c, d_src, pitch, N, M, out); break;
    default: launch_chunk_counts_g<32>(c, d_src, pitch, N, M, out); break;
  }
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// The per-SNP vectors of the rows a chunk keeps (index, mean, sd, 1/sd, mean/sd), compacted by the host into mapped
// pinned memory, are fetched by this kernel -- no small copies queue up in front of the next payload chunk.
__global__ void fetch_kept_kernel(const uint64_t* __restrict__ u_idx, const float* __restrict__ u_mean,
                                  const float* __restrict__ u_sd, const float* __restrict__ u_inv,
                                  const float* __restrict__ u_mu, uint64_t kept, uint64_t* __restrict__ d_idx,
                                  float* __restrict__ d_mean, float* __restrict__ d_sd, float* __restrict__ d_inv,
                                  float* __restrict__ d_mu) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < kept; t += (uint64_t)gridDim.x * blockDim.x) {
    d_idx[t] = u_idx[t];
    d_mean[t] = u_mean[t];
    d_sd[t] = u_sd[t];
    d_inv[t] = u_inv[t];
    d_mu[t] = u_mu[t];
  }
}

int launch_fetch_kept(gpca_ctx* c, const uint64_t* u_idx, const float* u_mean, const float* u_sd, const float* u_inv,
                      const float* u_mu, uint64_t kept, uint64_t* d_idx, float* d_mean, float* d_sd, float* d_inv,
                      float* d_mu) {
  if (kept == 0) return GPCA_OK;
  const int grid = (int)std::min<uint64_t>((kept + 255) / 256, (uint64_t)c->sm_count * 4);
  fetch_kept_kernel<<<grid, 256, 0, c->stream>>>(u_idx, u_mean, u_sd, u_inv, u_mu, kept, d_idx, d_mean, d_sd, d_inv, d_mu);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// Kept rows of the staged chunk -> rows [dst_row0, dst_row0 + kept) of the SNP-major resident matrix, recoded to
// dosage codes with zero pads.  d_idx holds loaded-row indices; the chunk starts at loaded row `row_base`.
// Destination rows at or past res_rows live in a ring of win_rows rows behind the resident part (GsLayout, common.cuh).
__global__ void build_gs_chunk_kernel(const uint8_t* __restrict__ src, size_t pitch, uint64_t N, uint64_t row_base,
                                      const uint64_t* __restrict__ idx, uint64_t kept, uint8_t* __restrict__ gs,
                                      size_t gs_pitch, uint64_t dst_row0, uint64_t res_rows, uint64_t win_rows) {
  const uint64_t chunks_per_row = gs_pitch / 16;
  const uint64_t total = kept * chunks_per_row;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w = t / chunks_per_row;
    const uint64_t ci = t - w * chunks_per_row;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ci * 64 < N) {
      v = load_row_chunk(src + (idx[w] - row_base) * pitch, N, ci);
      uint32_t x[4] = {recode_word(v.x), recode_word(v.y), recode_word(v.z), recode_word(v.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {      // (recoding turns the 00 pads into code 2: clear them again)
        const uint64_t k0 = ci * 64 + (uint64_t)j * 16;
        if (k0 >= N) x[j] = 0;
        else if (k0 + 16 > N) x[j] &= (1u << (2 * (int)(N - k0))) - 1u;
      }
      v = make_uint4(x[0], x[1], x[2], x[3]);
    }
    uint64_t drow = dst_row0 + w;
    if (drow >= res_rows && win_rows) drow = res_rows + (drow - res_rows) % win_rows;
    *reinterpret_cast<uint4*>(gs + drow * gs_pitch + ci * 16) = v;
  }
}

int launch_build_gs_chunk(gpca_ctx* c, const uint8_t* d_src, size_t pitch, uint64_t N, uint64_t row_base,
                          const uint64_t* d_idx, uint64_t kept, uint8_t* gs, size_t gs_pitch, uint64_t dst_row0,
                          uint64_t res_rows, uint64_t win_rows) {
  if (kept == 0) return GPCA_OK;
  const uint64_t total = kept * (gs_pitch / 16);
  const uint64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  build_gs_chunk_kernel<<<grid, 256, 0, c->stream>>>(d_src, pitch, N, row_base, d_idx, kept, gs, gs_pitch, dst_row0,
                                                     res_rows, win_rows);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// dst row i = src row idx[i] (idx < 0 -> zero row); 16-byte chunks
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, size_t src_pitch, const int64_t* __restrict__ idx,
                                   uint8_t* __restrict__ dst, size_t dst_pitch, uint64_t n_rows, size_t copy_bytes) {
  const uint64_t cpr = dst_pitch / 16;
  const uint64_t total = n_rows * cpr;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = t / cpr, ci = t - r * cpr;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int64_t s = idx[r];
    if (s >= 0 && ci * 16 < copy_bytes) v = ldg_nc_v4(src + (uint64_t)s * src_pitch + ci * 16);
    *reinterpret_cast<uint4*>(dst + r * dst_pitch + ci * 16) = v;
  }
}

int launch_gather_rows(gpca_ctx* c, PackedMat src, const int64_t* d_idx, PackedMat dst) {
  const uint64_t total = dst.rows * (dst.pitch / 16);
  if (!total) return GPCA_OK;
  const uint64_t blocks = (total + 255) / 256;
  const int grid = (int)std::min<uint64_t>(blocks, (uint64_t)c->sm_count * 16);
  gather_rows_kernel<<<grid, 256, 0, c->stream>>>(src.p, src.pitch, d_idx, dst.p, dst.pitch, dst.rows,
                                                  std::min(src.pitch, dst.pitch));
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// 2-bit transpose.  A CTA moves a tile of 512 source rows x 512 source fields: every thread transposes 16 x 16 blocks
// of 2-bit fields in registers (16 words in, 16 words out, four masked block-swap rounds), the 16 output words go to
// their destination rows in shared memory (XOR-swizzled columns: conflict-free both ways), and each destination row
// leaves as one 128-byte segment.  Reads are 128-byte segments per source row as well.  Every byte of the destination
// rows is written (pitches are multiples of 128 B), so no memset is needed.
__device__ __forceinline__ void transpose16x16_2bit(uint32_t (&x)[16]) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const uint32_t mask = (s == 8) ? 0x0000FFFFu : (s == 4) ? 0x00FF00FFu : (s == 2) ? 0x0F0F0F0Fu : 0x33333333u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i & s) continue;
      const uint32_t t = ((x[i] >> (2 * s)) ^ x[i + s]) & mask;   // upper-right block of row i <-> lower-left of row i+s
      x[i + s] ^= t;
      x[i] ^= t << (2 * s);
    }
  }
}

// block j (0..3) of a thread: source word column cw (16 fields), source row block rb (16 rows) of the 512 x 512 tile
__device__ __forceinline__ void tr_load16(const uint8_t* __restrict__ src, size_t src_pitch, uint64_t src_rows, uint64_t r0,
                                          uint64_t c0, int j, uint32_t (&x)[16]) {
  const int blk = threadIdx.x + 256 * j;
  const int cw = blk & 31, rb = blk >> 5;
  const uint8_t* p = src + (r0 + 16 * rb) * src_pitch + c0 / 4 + 4 * cw;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    x[i] = (r0 + 16 * rb + i < src_rows) ? __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)i * src_pitch)) : 0u;
}
__device__ __forceinline__ void tr_store16(uint32_t* tile, int j, uint32_t (&x)[16]) {
  const int blk = threadIdx.x + 256 * j;
  const int cw = blk & 31, rb = blk >> 5;
  transpose16x16_2bit(x);
#pragma unroll
  for (int i = 0; i < 16; ++i) tile[(16 * cw + i) * 32 + (rb ^ cw)] = x[i];
}

__global__ void __launch_bounds__(256, 3) transpose2bit_kernel(const uint8_t* __restrict__ src, size_t src_pitch,
                                                            uint64_t src_rows, uint64_t src_cols,
                                                            uint8_t* __restrict__ dst, size_t dst_pitch) {
  extern __shared__ uint32_t tile[];   // [512 destination rows][32 words]
  const uint64_t tiles_c = (src_cols + 511) / 512;
  const uint64_t tiles_r = (src_rows + 511) / 512;
  // Tile order: blocks of TB_R x TB_C tiles, row tiles fastest inside a block.  The CTAs that run at the same time
  // (a few hundred) then read TB_C adjacent 128-byte segments of every source row and write TB_R adjacent segments of
  // every destination row -- kilobytes per DRAM page on both sides.  (With one column of tiles after the other, every
  // 128-byte read opened its own DRAM page when the source pitch is tens of KB: 1.45 TB/s.)
  constexpr uint64_t TB_R = 32, TB_C = 16;
  const uint64_t blocks_r = (tiles_r + TB_R - 1) / TB_R, blocks_c = (tiles_c + TB_C - 1) / TB_C;
  const uint64_t ntiles = blocks_r * blocks_c * TB_R * TB_C;
  for (uint64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const uint64_t blk = tix / (TB_R * TB_C), in = tix - blk * (TB_R * TB_C);
    const uint64_t bc = blk / blocks_r, br = blk - bc * blocks_r;
    const uint64_t tr = br * TB_R + in % TB_R, tc = bc * TB_C + in / TB_R;
    if (tr >= tiles_r || tc >= tiles_c) continue;      // (uniform per CTA: no barrier is skipped by part of a CTA)
    const uint64_t r0 = tr * 512, c0 = tc * 512;
    // Four 16 x 16 blocks per thread, software-pipelined: the 16 loads of block j + 1 are issued before block j is
    // transposed and stored to shared memory, so every thread keeps global loads in flight while it computes.  (With
    // one block at a time the kernel alternated between a load phase and a compute phase: 24 warps per SM with nothing
    // outstanding two thirds of the time -- 1.36 TB/s, 16 % DRAM throughput, 8.9 long-scoreboard stalls per issue.)
    uint32_t xa[16], xb[16];
    tr_load16(src, src_pitch, src_rows, r0, c0, 0, xa);
    tr_load16(src, src_pitch, src_rows, r0, c0, 1, xb);
    tr_store16(tile, 0, xa);
    tr_load16(src, src_pitch, src_rows, r0, c0, 2, xa);
    tr_store16(tile, 1, xb);
    tr_load16(src, src_pitch, src_rows, r0, c0, 3, xb);
    tr_store16(tile, 2, xa);
    tr_store16(tile, 3, xb);
    __syncthreads();
    // destination row dr (= source column c0 + dr): 32 words, word j holds source rows r0 + 16 j .. + 15
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int dr = warp; dr < 512; dr += 8) {
      const uint64_t dn = c0 + dr;
      if (dn >= src_cols) break;
      const uint32_t v = tile[dr * 32 + (lane ^ ((dr >> 4) & 31))];
      reinterpret_cast<uint32_t*>(dst + dn * dst_pitch + r0 / 4)[lane] = v;
    }
    __syncthreads();
  }
}

int launch_transpose(gpca_ctx* c, PackedMat gs, PackedMat gt) {
  if (gs.rows == 0 || gs.cols == 0) return GPCA_OK;
  // the source pitch covers whole 128-byte column tiles and the destination pitch covers round_up(src_rows, 512) / 4
  // bytes: both hold because every pitch is a multiple of 128 B
  if (gs.pitch % 128 != 0 || gt.pitch % 128 != 0 || gt.pitch * 4 < gs.rows) {
    c->set_error("transpose: pitches must be multiples of 128 bytes");
    return GPCA_ERR_INVALID;
  }
  const uint64_t tiles_c = (gs.cols + 511) / 512, tiles_r = (gs.rows + 511) / 512;
  const uint64_t ntiles = ((tiles_r + 31) / 32) * ((tiles_c + 15) / 16) * 512;      // padded to whole blocks of tiles
  const int grid = (int)(ntiles < (uint64_t)c->sm_count * 3 ? ntiles : (uint64_t)c->sm_count * 3);
  const int smem = 512 * 32 * 4;
  GPCA_CUDA_TRY(c, cudaFuncSetAttribute(transpose2bit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  transpose2bit_kernel<<<grid, 256, smem, c->stream>>>(gs.p, gs.pitch, gs.rows, gs.cols, gt.p, gt.pitch);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Accessor parity kernel (prepare.rs:1884-2016): f32 fma standardisation of a gathered block.
// SNP rows at or past gs_rows are not resident in the SNP-major matrix (gpca_ctx::gs_res_rows): their fields are read
// from the sample-major one.
__global__ void std_block_kernel(const uint8_t* __restrict__ gs, size_t pitch, uint64_t gs_rows,
                                 const uint8_t* __restrict__ gt, size_t gt_pitch, const float* __restrict__ mean,
                                 const float* __restrict__ sd, const uint64_t* __restrict__ ids, uint64_t n_ids,
                                 const uint64_t* __restrict__ samp, uint64_t n_samp, float* __restrict__ out,
                                 int* __restrict__ missing_flag) {
  const uint64_t total = n_ids * n_samp;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = t / n_samp, j = t - i * n_samp;
    const uint64_t id = ids[i];
    const uint64_t s = samp ? samp[j] : j;
    const uint32_t code = id < gs_rows ? (gs[id * pitch + (s >> 2)] >> (2 * (s & 3))) & 3u
                                       : (gt[s * gt_pitch + (id >> 2)] >> (2 * (id & 3))) & 3u;
    const float m = mean[id], sdev = sd[id];
    float z = 0.0f;
    if (code == 3u) {
      atomicExch(missing_flag, 1);
    } else if (!(fabsf(sdev) < 1e-9f)) {
      const float recip = __fdiv_rn(1.0f, sdev);
      const float bterm = __fmul_rn(-m, recip);
      z = __fmaf_rn((float)code, recip, bterm);
    }
    out[t] = z;
  }
}

int launch_std_block(gpca_ctx* c, PackedMat gs, uint64_t gs_rows, PackedMat gt, const float* d_mean, const float* d_sd, const uint64_t* d_ids,
                     uint64_t n_ids, const uint64_t* d_samp, uint64_t n_samp, float* d_out, int* d_missing_flag) {
  const uint64_t total = n_ids * n_samp;
  if (total == 0) return GPCA_OK;
  const int threads = 256;
  const uint64_t blocks = (total + threads - 1) / threads;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  std_block_kernel<<<grid, threads, 0, c->stream>>>(gs.p, gs.pitch, gs_rows, gt.p, gt.pitch, d_mean, d_sd, d_ids, n_ids, d_samp, n_samp, d_out,
                                                    d_missing_flag);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}

// ------------------------------------------------------------------------------------------
// Synthetic structured genotypes for benchmarks, written straight in PLINK .bed layout.
// Counter-based: every value is a function of (seed, global SNP index, sample index) only, so any shard of SNPs
// regenerates identically at any GPU count.  Balding-Nichols with the normal approximation of the beta:
//   ancestral frequency p ~ U(0.05, 0.5), population frequency p_k = clamp(p + sqrt(F p (1-p)) z_k, 0.01, 0.99),
//   genotype = Binomial(2, p_pop(sample)), optional missing calls (code 01) at `missing_rate`.
// One thread per output byte (4 samples): two Philox calls -> 8 uniforms (2 per sample) + one for missingness.
constexpr uint32_t SYNTH_STREAM_FREQ = 0x5EED0001u, SYNTH_STREAM_GENO = 0x5EED0002u, SYNTH_STREAM_MISS = 0x5EED0003u;

// fst_grade > 0 grades the drift of the populations: F_ST of population k = fst * (1 + grade * (0.5 - k / (P - 1))),
// so that the P - 1 structural eigenvalues of the standardized matrix are distinct (with equal populations and one
// F_ST they form one degenerate cluster, and a top-k subspace that cuts through it is not defined).
__device__ __forceinline__ float synth_pop_fst(uint32_t pop, uint32_t n_pops, float fst, float grade) {
  if (grade == 0.0f || n_pops < 2) return fst;
  return fst * (1.0f + grade * (0.5f - (float)pop / (float)(n_pops - 1)));
}
__device__ __forceinline__ float synth_pop_freq(uint64_t seed, uint64_t snp, uint32_t pop, float fst) {
  const Philox4 r = philox4x32_10((uint32_t)snp, (uint32_t)(snp >> 32), 0u, SYNTH_STREAM_FREQ, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  const float p = 0.05f + 0.45f * ((float)(r.x >> 8) * (1.0f / 16777216.0f));
  const float z = philox_normal(seed, SYNTH_STREAM_FREQ, snp, 4u + pop);   // columns 4.. : population deviates
  const float q = p + sqrtf(fst * p * (1.0f - p)) * z;
  return fminf(fmaxf(q, 0.01f), 0.99f);
}

__global__ void __launch_bounds__(256) synth_bed_kernel(uint8_t* __restrict__ out, uint64_t n_samples, uint64_t n_snps,
                                                        uint64_t snp_offset, uint64_t seed, uint32_t n_pops, float fst,
                                                        float missing_rate, float fst_grade) {
  const uint64_t bps = (n_samples + 3) / 4;
  const uint64_t total = n_snps * bps;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = t / bps, b = t - j * bps;
    const uint64_t snp = snp_offset + j;
    const Philox4 u0 = philox4x32_10((uint32_t)snp, (uint32_t)(snp >> 32), (uint32_t)(2 * b), SYNTH_STREAM_GENO,
                                     (uint32_t)seed, (uint32_t)(seed >> 32));
    const Philox4 u1 = philox4x32_10((uint32_t)snp, (uint32_t)(snp >> 32), (uint32_t)(2 * b + 1), SYNTH_STREAM_GENO,
                                     (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    Philox4 um = {0, 0, 0, 0};
    if (missing_rate > 0.0f)
      um = philox4x32_10((uint32_t)snp, (uint32_t)(snp >> 32), (uint32_t)b, SYNTH_STREAM_MISS, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
    const uint32_t umv[4] = {um.x, um.y, um.z, um.w};
    uint32_t byte = 0;
    uint32_t last_pop = 0xffffffffu;
    float pf = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t i = 4 * b + q;
      if (i >= n_samples) break;                       // pad fields stay 00 (as in a real .bed)
      const uint32_t pop = (uint32_t)(i * n_pops / n_samples);
      if (pop != last_pop) {
        pf = synth_pop_freq(seed, snp, pop, synth_pop_fst(pop, n_pops, fst, fst_grade));
        last_pop = pop;
      }
      const uint32_t thr = (uint32_t)(pf * 16777216.0f);
      const uint32_t dosage = ((u[2 * q] >> 8) < thr ? 1u : 0u) + ((u[2 * q + 1] >> 8) < thr ? 1u : 0u);
      uint32_t code = dosage == 0 ? 3u : (dosage == 1 ? 2u : 0u);    // count_a1: 0 -> 11, 1 -> 10, 2 -> 00
      if (missing_rate > 0.0f && (float)(umv[q] >> 8) * (1.0f / 16777216.0f) < missing_rate) code = 1u;
      byte |= code << (2 * q);
    }
    out[t] = (uint8_t)byte;
  }
}

int launch_synth_bed(gpca_ctx* c, uint8_t* d_out, uint64_t n_samples, uint64_t n_snps, uint64_t snp_offset, uint64_t seed,
                     uint32_t n_pops, float fst, float missing_rate, float fst_grade) {
  const uint64_t total = n_snps * ((n_samples + 3) / 4);
  if (total == 0) return GPCA_OK;
  const uint64_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < (uint64_t)c->sm_count * 16 ? blocks : (uint64_t)c->sm_count * 16);
  synth_bed_kernel<<<grid, 256, 0, c->stream>>>(d_out, n_samples, n_snps, snp_offset, seed, n_pops ? n_pops : 1, fst,
                                                missing_rate, fst_grade);
  KLAUNCH_CHECK(c);
  return GPCA_OK;
}
