// FP64 / FP32 FMA and conversion throughput per SM (measurement probe; prints ops/clk/SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const float fa = (float)a, fb = (float)b;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else if (MODE == 1) {
      f0 = fmaf(f0, fa, fb); f1 = fmaf(f1, fa, fb); f2 = fmaf(f2, fa, fb); f3 = fmaf(f3, fa, fb);
      f4 = fmaf(f4, fa, fb); f5 = fmaf(f5, fa, fb); f6 = fmaf(f6, fa, fb); f7 = fmaf(f7, fa, fb);
    } else {   // f32 -> f64 conversion chained through an fp32 op
      x0 = (double)f0; f0 = (float)x0 + fa; x1 = (double)f1; f1 = (float)x1 + fa;
      x2 = (double)f2; f2 = (float)x2 + fa; x3 = (double)f3; f3 = (float)x3 + fa;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sms * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms * 4, 1024>>>(out, iters, 1.0000001, 1e-9);
      if (mode == 1) k<1><<<sms * 4, 1024>>>(out, iters, 1.0000001, 1e-9);
      if (mode == 2) k<2><<<sms * 4, 1024>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * 4 * 1024 * iters * (mode == 2 ? 8.0 : 8.0);
    printf("%s: %.3f ms, %.2f Tops/s, %.1f ops/clk/SM at the nominal %d MHz\n", mode == 0 ? "DFMA" : mode == 1 ? "FFMA" : "F2F (f32<->f64 conversions)",
           ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  }
  return 0;
}
