// Shared device helpers of libtqf: error plumbing, Philox4x32-10, the
// TensorFlow uint->float conversions, Box-Muller, inverse normal CDF and the
// Sobol XOR machinery.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <string>

#include "tqf.h"
#include "tqf_math.cuh"

namespace tqf {

// ------------------------------------------------------------ errors ------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
// Device-global {T, 1/c} table of the table logarithm (tqf_rng.cu).
int device_logtab(const double** out);
// Device-global cubic table of the float32 inverse normal CDF (tqf_ndtri_f32_tab.inc).
int device_ndtri_f32_tab(const float** out);
// Device buffers of plans (coefficient tables, direction numbers, partial sums): served
// from a per-device free list of power-of-two blocks, so that a pricing call with new
// parameters -- a new plan -- costs no cudaMalloc / cudaFree once the process is warm
// (tqf_rng.cu).  dev_free_all_sync(): one device synchronisation, then every block goes
// back to the list (the semantics cudaFree had: nothing in flight can still read them).
int dev_alloc(void** out, size_t bytes);
void dev_release(void* const* ptrs, int count);

#define TQF_CUDA_OK(expr)                                        \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    if (e__ != cudaSuccess) return ::tqf::cuda_fail(e__, #expr); \
  } while (0)

#define TQF_REQUIRE(cond, msg)            \
  do {                                    \
    if (!(cond)) {                        \
      ::tqf::set_error(msg);              \
      return TQF_ERR_INVALID_ARGUMENT;    \
    }                                     \
  } while (0)

// NVTX range around a C-ABI entry point (header-only NVTX 3: a no-op function
// pointer check unless a profiler injected itself into the process).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define TQF_NVTX(name) ::tqf::NvtxRange nvtx_range__(name)

constexpr int kSMs = 148;  // B200

// ------------------------------------------------------------ Philox ------
// Philox4x32-10 exactly as Random123 / tensorflow/core/lib/random/philox_random.h.
struct PhiloxKey {
  uint32_t k0, k1;
};
struct PhiloxCtr {
  uint32_t c0, c1, c2, c3;
};

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1,
                                                         uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 128-bit counter + 64-bit group number (PhiloxRandom::Skip).
__host__ __device__ __forceinline__ uint4 philox_group(const PhiloxCtr& base,
                                                        const PhiloxKey& key,
                                                        uint64_t g) {
  const uint64_t base_lo = (static_cast<uint64_t>(base.c1) << 32) | base.c0;
  const uint64_t base_hi = (static_cast<uint64_t>(base.c3) << 32) | base.c2;
  const uint64_t lo = base_lo + g;
  const uint64_t hi = base_hi + (lo < base_lo ? 1ull : 0ull);
  return philox4x32_10(static_cast<uint32_t>(lo), static_cast<uint32_t>(lo >> 32),
                       static_cast<uint32_t>(hi), static_cast<uint32_t>(hi >> 32),
                       key.k0, key.k1);
}

// The same for kernels that carry `lo` = low 64 counter bits + group number themselves
// (one 64-bit increment per refill instead of a 128-bit add with carry detection: 7
// instructions per Philox call).  Valid while the low half cannot wrap: plans require a
// base below 2^63 (TensorFlow's seed derivations start it at 0), groups stay below 2^62.
__device__ __forceinline__ uint64_t philox_base_lo(const PhiloxCtr& base) {
  return (static_cast<uint64_t>(base.c1) << 32) | base.c0;
}
__device__ __forceinline__ uint4 philox_group_nowrap(const PhiloxCtr& base, const PhiloxKey& key,
                                                     uint64_t lo) {
  return philox4x32_10(static_cast<uint32_t>(lo), static_cast<uint32_t>(lo >> 32), base.c2, base.c3,
                       key.k0, key.k1);
}

// tensorflow/core/lib/random/random_distributions.h: Uint64ToDouble.
__device__ __forceinline__ double uint64_to_double(uint32_t x0, uint32_t x1) {
  const uint32_t hi = (x0 & 0xFFFFFu) | 0x3FF00000u;
  return __hiloint2double(static_cast<int>(hi), static_cast<int>(x1)) - 1.0;
}
// ... Uint32ToFloat.
__device__ __forceinline__ float uint32_to_float(uint32_t x) {
  return __uint_as_float((x & 0x7FFFFFu) | 0x3F800000u) - 1.0f;
}

// BoxMullerDouble: (x0,x1,x2,x3) -> (r sin v, r cos v).
__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, uint32_t x2,
                                           uint32_t x3, double* n0, double* n1) {
  double u1 = uint64_to_double(x0, x1);
  u1 = u1 < 1.0e-7 ? 1.0e-7 : u1;
  const double v1 = 6.283185307179586476925286766559 * uint64_to_double(x2, x3);
  const double u2 = fm::sqrt_pos(-2.0 * fm::log_pos(u1));
  double s, c;
  fm::sincos_2pi(v1, &s, &c);
  *n0 = s * u2;
  *n1 = c * u2;
}
// BoxMullerFloat.
__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, float* n0,
                                           float* n1) {
  float u1 = uint32_to_float(x0);
  u1 = u1 < 1.0e-7f ? 1.0e-7f : u1;
  // TF: `2.0f * M_PI * Uint32ToFloat(x1)` -- the product is formed in double.
  const float v1 = static_cast<float>(
      6.283185307179586476925286766559 * static_cast<double>(uint32_to_float(x1)));
  const float u2 = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(v1, &s, &c);
  *n0 = s * u2;
  *n1 = c * u2;
}

// --------------------------------------------------------- inverse CDF ----
// sqrt(2) erfinv(2u - 1) == ndtri(u) (multivariate_normal.py:420).
// `logtab`: device-global table of the table logarithm (tqf::device_logtab()).
__device__ __forceinline__ double ndtri(double u, const double* logtab) {
  return fm::ndtri_q(u - 0.5, logtab);
}
// (u - 0.5) * 2 as the reference forms it (multivariate_normal.py:420), exact in fp32.
__device__ __forceinline__ float ndtri(float u) { return fm::ndtri_t_f32((u - 0.5f) * 2.0f); }
__device__ __forceinline__ float ndtri(float u, const double*) { return ndtri(u); }

// ------------------------------------------------------------- Sobol ------
// Device table V[d][32]: direction number m[d][b] left-aligned in 32 bits,
// V[d][b] = m[d][b] << (31 - b).  A point is x32 = XOR_b bit_b(i) V[d][b] and
// its value is x32 / 2^32 -- the same dyadic rational as the reference's
// x / 2^num_digits (sobol_impl.py:128-167) for every num_digits.
__device__ __forceinline__ double sobol_uniform_f64(uint32_t x32) {
  // (2^52 + x32) * 2^-32 - 2^20, exact.
  return __hiloint2double(0x41300000, static_cast<int>(x32)) - 1048576.0;
}
// t = 2u - 1 = x32 / 2^31 - 1 in ONE exact subtraction: (2^21 + x32 2^-31) - (2^21 + 1).
__device__ __forceinline__ double sobol_centered_f64(uint32_t x32) {
  return __hiloint2double(0x41400000, static_cast<int>(x32)) - 2097153.0;
}
__device__ __forceinline__ float sobol_uniform_f32(uint32_t x32) {
  // The reference casts the integer point to float32 (round to nearest even)
  // and divides by a power of two: identical to rounding x32 and scaling.
  return __uint2float_rn(x32) * 2.3283064365386963e-10f;
}

template <typename Real>
struct RealTraits;
template <>
struct RealTraits<double> {
  static constexpr int kDtype = TQF_F64;
  static constexpr int kPerGroup = 2;  // normals per Philox call
  __device__ static __forceinline__ double sobol_uniform(uint32_t x) {
    return sobol_uniform_f64(x);
  }
};
template <>
struct RealTraits<float> {
  static constexpr int kDtype = TQF_F32;
  static constexpr int kPerGroup = 4;
  __device__ static __forceinline__ float sobol_uniform(uint32_t x) {
    return sobol_uniform_f32(x);
  }
};

// ----------------------------------------------------------- reductions ---
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

}  // namespace tqf
