// dense_tc.cuh -- products with the dense fp32 condensed-feature matrix (EigenSNP global rSVD), see dense_tc.cu
#pragma once
#include "kernels.cuh"

// out[r, :] = a_r * sum_k X[r, k] f_k Bin[k, :]  -  b_r * sum_k e_k Bin[k, :]
//   cols_mode = false: X = C   [N x R]  (rows = samples,  k = condensed features)
//   cols_mode = true : X = C^T [R x N]  (rows = features, k = samples)
struct DenseProduct {
  const float* C;        // [N x R] row-major, row stride ldc floats (ldc % 4 == 0 for the tensor path)
  uint64_t N, R;
  uint32_t ldc;
  bool cols_mode;
  const float* Bin;      // [K x ld]
  uint32_t l, ld;
  const float* f;        // [K] or null (= 1)
  const float* e;        // [K] or null (= 1)
  const float* a;        // [rows] or null (= 1)
  const float* b;        // [rows] or null (= 1)
  float* out;            // [rows x ldo]
  uint32_t ldo;
};
int launch_dense_product(gpca_ctx* c, const DenseProduct& p);
