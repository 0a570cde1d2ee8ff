// philox.cuh -- Philox4x32-10 counter-based RNG and a Box-Muller normal.
// Twin of oracle/rng.py (tests pin both against the Random123 known-answer vectors).
// Stands in for the seeded ChaCha stream inside efficient_pca (seed plumbed at
// src/main.rs:637 `--rfit-seed` and src/main.rs:321 `--eigensnp-seed`).
#pragma once
#include <stdint.h>

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// (|z| of any generated normal is below GPCA_NORMAL_ABS_MAX, kernels.cuh: u1 >= 2^-25, so the Box-Muller radius is at
//  most sqrt(50 ln 2) = 5.887)
// Four standard normals for columns 4*cg .. 4*cg+3 of row `row` (one Philox call, two Box-Muller pairs):
//   (x, y) -> sqrt(-2 ln u1) * {cos, sin}(2 pi u2),   (z, w) -> the same for the second pair.
__device__ inline void philox_normal4(uint64_t seed, uint32_t stream, uint64_t row, uint32_t cg, float out[4]) {
  const Philox4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), cg, stream, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  const float u1a = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2a = (float)(r.y >> 8) * (1.0f / 16777216.0f);
  const float u1b = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2b = (float)(r.w >> 8) * (1.0f / 16777216.0f);
#ifndef GPCA_FAST_NORMAL
#define GPCA_FAST_NORMAL 1
#endif
#if GPCA_FAST_NORMAL
  // hardware approximations (MUFU lg2 / sin / cos, about 2^-21 absolute): the test matrices only have to be Gaussian
  // enough for a randomized range finder, and agree with the oracle's float32 twin to ~1e-6
  const float ra = sqrtf(-2.0f * __logf(u1a)), rb = sqrtf(-2.0f * __logf(u1b));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u2a, &sa, &ca);
  __sincosf(6.283185307179586f * u2b, &sb, &cb);
#else
  const float ra = sqrtf(-2.0f * logf(u1a)), rb = sqrtf(-2.0f * logf(u1b));
  float sa, ca, sb, cb;
  sincospif(2.0f * u2a, &sa, &ca);
  sincospif(2.0f * u2b, &sb, &cb);
#endif
  out[0] = ra * ca;
  out[1] = ra * sa;
  out[2] = rb * cb;
  out[3] = rb * sb;
}

// One standard normal for element (row, col) of stream `stream` under `seed` (column col = 4*cg + j).
__device__ inline float philox_normal(uint64_t seed, uint32_t stream, uint64_t row, uint32_t col) {
  float o[4];
  philox_normal4(seed, stream, row, col >> 2, o);
  return o[col & 3];
}
