// DFMA throughput microbenchmarks for the shapes the path kernel issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__constant__ __align__(16) double g_c[32] = {1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32};

// MODE 3: LDS.128 double-buffered one pair ahead.
template <int K>
__global__ void __launch_bounds__(128) horner_prefetch_kernel(double* out, int iters, double seed) {
  __shared__ __align__(16) double s_c[32];
  if (threadIdx.x < 32) s_c[threadIdx.x] = 1.0 / (1.0 + threadIdx.x);
  __syncthreads();
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(s_c));
  double y[K], p[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { y[k] = seed + 1e-3 * (threadIdx.x + k); p[k] = y[k]; }
  for (int it = 0; it < iters; ++it) {
    double c0, c1, n0, n1;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(base));
#pragma unroll
    for (int i = 0; i < 24; i += 2) {
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n0), "=d"(n1) : "r"(base + ((i + 2) % 24) * 8u));
#pragma unroll
      for (int k = 0; k < K; ++k) p[k] = fma(fma(p[k], y[k], c0), y[k], c1);
      c0 = n0; c1 = n1;
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) s += p[k];
  if (s == -1.2345) out[0] = s;
}

template <int K, int MODE>
__global__ void __launch_bounds__(128) horner_kernel(double* out, int iters, double seed) {
  __shared__ __align__(16) double s_c[32];
  if (threadIdx.x < 32) s_c[threadIdx.x] = 1.0 / (1.0 + threadIdx.x);
  __syncthreads();
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(s_c));
  double y[K], p[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { y[k] = seed + 1e-3 * (threadIdx.x + k); p[k] = y[k]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 24; i += 2) {
      double c0, c1;
      if (MODE == 0) {           // coefficient pair from volatile LDS.128
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(base + i * 8u));
      } else if (MODE == 1) {    // immediates folded by the compiler (uniform regs)
        c0 = 0.123456789 + i; c1 = 0.987654321 - i;
      } else if (MODE == 2) {    // c depends on registers only (no load)
        c0 = y[0]; c1 = y[K - 1];
      } else if (MODE == 4) {    // volatile constant-bank load
        asm volatile("ld.const.v2.f64 {%0, %1}, [%2];" : "=d"(c0), "=d"(c1) : "l"(__cvta_generic_to_constant(g_c + i)));
      } else {
        c0 = 0; c1 = 0;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) p[k] = fma(fma(p[k], y[k], c0), y[k], c1);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) s += p[k];
  if (s == -1.2345) out[0] = s;
}

template <int K, int MODE>
double run(int blocks_per_sm) {
  double* out; cudaMalloc(&out, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * blocks_per_sm, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    if (MODE == 3) horner_prefetch_kernel<K><<<grid, 128>>>(out, iters, 0.5);
    else horner_kernel<K, MODE><<<grid, 128>>>(out, iters, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = double(grid) * 128 * iters * 24.0 * K / (ms * 1e-3);
    if (r && rate > best) best = rate;
  }
  cudaFree(out);
  return best;
}

// Dispatch-port test: K = 8 DFMA chains with NI independent integer ops (LOP3)
// after every ND DFMAs.  If a DFMA held the SMSP dispatch port for its 2 pipe
// cycles, the integer ops would add to the run time; if they issue in the
// shadow of the FP64 pipe, the DFMA rate stays at the peak.
template <int NI, int ND>
__global__ void __launch_bounds__(128) mix_kernel(double* out, int iters, double seed) {
  double y[8], p[8];
  uint32_t q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { y[k] = seed + 1e-3 * (threadIdx.x + k); p[k] = y[k]; q[k] = threadIdx.x * 7 + k; }
  const double c = seed * 0.25;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 24; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(p[k]) : "d"(y[k]), "d"(c));
        if (((i * 8 + k) % ND) == ND - 1) {
#pragma unroll
          for (int n = 0; n < NI; ++n)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[(k + n) & 7]) : "r"(q[(k + n + 3) & 7]), "r"(it));
        }
      }
    }
  }
  double s = 0;
  uint32_t t = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { s += p[k]; t ^= q[k]; }
  if (s == -1.2345 || t == 0x12345u) out[0] = s + t;
}

template <int NI, int ND>
double run_mix(int blocks_per_sm) {
  double* out; cudaMalloc(&out, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * blocks_per_sm, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    mix_kernel<NI, ND><<<grid, 128>>>(out, iters, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = double(grid) * 128 * iters * 24.0 * 8 / (ms * 1e-3);
    if (r && rate > best) best = rate;
  }
  cudaFree(out);
  return best;
}

int main() {
  printf("dispatch test (DFMA/s): int ops per DFMA\n");
  for (int b : {3, 4}) {
    printf("b=%d  0    %.3e\n", b, run_mix<0, 1>(b));
    printf("b=%d  1/4  %.3e\n", b, run_mix<1, 4>(b));
    printf("b=%d  1/2  %.3e\n", b, run_mix<1, 2>(b));
    printf("b=%d  1    %.3e\n", b, run_mix<1, 1>(b));
    printf("b=%d  2    %.3e\n", b, run_mix<2, 1>(b));
  }
  printf("K MODE blocks/SM  DFMA/s\n");
  for (int b : {4}) {
    printf("8 pref  %d %.3e\n", b, run<8, 3>(b));
    printf("4 pref  %d %.3e\n", b, run<4, 3>(b));
    printf("8 ldc   %d %.3e\n", b, run<8, 4>(b));
    printf("4 ldc   %d %.3e\n", b, run<4, 4>(b));
    printf("4 lds   %d %.3e\n", b, run<4, 0>(b));
    printf("4 imm   %d %.3e\n", b, run<4, 1>(b));
    printf("4 reg   %d %.3e\n", b, run<4, 2>(b));
    printf("8 lds   %d %.3e\n", b, run<8, 0>(b));
    printf("8 imm   %d %.3e\n", b, run<8, 1>(b));
    printf("2 lds   %d %.3e\n", b, run<2, 0>(b));
    printf("1 lds   %d %.3e\n", b, run<1, 0>(b));
  }
  return 0;
}
