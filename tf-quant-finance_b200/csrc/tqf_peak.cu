// FP64 / FP32 FMA issue-rate microbenchmark: the roofline denominator of the
// fused (compute-bound) mode.  MEASURED_PEAKS.json carries HBM and bf16 tensor
// peaks only, so the DFMA peak is measured live on the device the bench uses.
#include "tqf_common.cuh"

namespace tqf {

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5,
    r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      r0 = fma(r0, a, b);
      r1 = fma(r1, a, b);
      r2 = fma(r2, a, b);
      r3 = fma(r3, a, b);
      r4 = fma(r4, a, b);
      r5 = fma(r5, a, b);
      r6 = fma(r6, a, b);
      r7 = fma(r7, a, b);
    }
  }
  const T s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
  if (s == T(-1.2345)) out[0] = s;  // never true; keeps the chains alive
}

// Second shape: 8 interleaved Horner chains with compile-time coefficients
// (the shape of the kernels' polynomial evaluation).  The reported peak is the
// best of the two shapes.
template <typename T>
__global__ void __launch_bounds__(128) fma_peak_horner_kernel(T* out, int iters, T seed) {
  T y[8], p[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    y[k] = seed + T(1e-3) * T(threadIdx.x + k);
    p[k] = y[k];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const T c0 = T(0.123456789) + T(i), c1 = T(0.987654321) - T(i);
#pragma unroll
      for (int k = 0; k < 8; ++k) p[k] = fma(fma(p[k], y[k], c0), y[k], c1);
    }
  }
  T s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += p[k];
  if (s == T(-1.2345)) out[0] = s;
}

template <typename T>
static int measure(double* per_second) {
  T* out = nullptr;
  TQF_CUDA_OK(cudaMalloc(&out, sizeof(T)));
  int sms = kSMs;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, block = 256, iters = 4096;
  cudaEvent_t e0, e1;
  TQF_CUDA_OK(cudaEventCreate(&e0));
  TQF_CUDA_OK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    TQF_CUDA_OK(cudaEventRecord(e0));
    fma_peak_kernel<T><<<grid, block>>>(out, iters, T(0.999999), T(1e-6));
    TQF_CUDA_OK(cudaEventRecord(e1));
    TQF_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    TQF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = static_cast<double>(grid) * block * iters * 64.0;
    const double rate = fmas / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  for (int rep = 0; rep < 5; ++rep) {
    const int hgrid = sms * 8, hiters = 1024;
    TQF_CUDA_OK(cudaEventRecord(e0));
    fma_peak_horner_kernel<T><<<hgrid, 128>>>(out, hiters, T(0.5));
    TQF_CUDA_OK(cudaEventRecord(e1));
    TQF_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    TQF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = static_cast<double>(hgrid) * 128 * hiters * 32.0 * 8.0;
    const double rate = fmas / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *per_second = best;
  return TQF_OK;
}

}  // namespace tqf

extern "C" int tqf_measure_fp64_peak(double* dfma_per_second, double* ffma_per_second) {
  TQF_REQUIRE(dfma_per_second && ffma_per_second, "null argument");
  int rc = tqf::measure<double>(dfma_per_second);
  if (rc != TQF_OK) return rc;
  return tqf::measure<float>(ffma_per_second);
}
