// tff.math.qmc on the device: digital nets (Sobol generating matrices, linear
// matrix scrambling, digital shift), rank-1 lattice rules and TensorFlow's
// stateless integer uniform that seeds the randomisations.
//
// Reference: math/qmc/digital_net.py:45-527, math/qmc/sobol.py:32-395,
// math/qmc/lattice_rule.py:40-229, math/qmc/utils.py:23-158.
// The [dim, log2 n] integer tables are built on the host (they are a few KB);
// the points are produced by HBM-write-bound kernels.  sm_100a only.
#include <vector>

#include "tqf_common.cuh"

namespace tqf {
namespace {

int qmc_grid(uint64_t work_items, int block) {
  uint64_t blocks = (work_items + block - 1) / block;
  const uint64_t cap = static_cast<uint64_t>(kSMs) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return static_cast<int>(blocks);
}

// UniformDistribution<PhiloxRandom, int32 / int64>
// (tensorflow/core/lib/random/random_distributions.h): lo + x % range.
__global__ void philox_uniform_int32_kernel(PhiloxKey key, PhiloxCtr ctr, int32_t lo, uint32_t range,
                                            uint64_t n, int32_t* __restrict__ out) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t groups = (n + 3) / 4;
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups;
       g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t e = 4 * g + j;
      if (e < n) out[e] = static_cast<int32_t>(static_cast<uint32_t>(lo) + x[j] % range);
    }
  }
}

__global__ void philox_uniform_int64_kernel(PhiloxKey key, PhiloxCtr ctr, int64_t lo, uint64_t range,
                                            uint64_t n, int64_t* __restrict__ out) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t groups = (n + 1) / 2;
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups;
       g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    const uint64_t x0 = (static_cast<uint64_t>(w.y) << 32) | w.x;
    const uint64_t x1 = (static_cast<uint64_t>(w.w) << 32) | w.z;
    if (2 * g < n) out[2 * g] = static_cast<int64_t>(static_cast<uint64_t>(lo) + x0 % range);
    if (2 * g + 1 < n) out[2 * g + 1] = static_cast<int64_t>(static_cast<uint64_t>(lo) + x1 % range);
  }
}

template <typename Real>
__device__ __forceinline__ Real tent(Real v) {
  // utils.py:94-117: where(v < 0.5, 2 v, 2 (1 - v))
  return v < Real(0.5) ? Real(2) * v : Real(2) * (Real(1) - v);
}

// One element (point i, coordinate d) per thread and grid-stride step; d runs
// fastest, so a warp writes 32 consecutive coordinates of (mostly) one point and
// the bit tests of the index are warp-uniform.  gt is the TRANSPOSED table
// [m][dim] (coalesced, L1-resident).  Int is the reference's int_dtype: the
// integer -> real cast and the divisor 1 << num_digits follow its wrap-around.
template <typename Int, typename Real>
__global__ void digital_net_kernel(const Int* __restrict__ gt, const Int* __restrict__ shift,
                                   const int64_t* __restrict__ seq, uint64_t first_index,
                                   uint64_t count, int dim, int m, Real denom, int apply_tent,
                                   Real* __restrict__ out) {
  const uint64_t total = count * static_cast<uint64_t>(dim);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += stride) {
    const uint64_t i = e / dim;
    const int d = static_cast<int>(e - i * dim);
    const uint64_t idx = seq ? static_cast<uint64_t>(seq[i]) : first_index + i;
    Int x = shift[d];
    for (int b = 0; b < m; ++b) {
      if ((idx >> b) & 1ull) x ^= __ldg(gt + static_cast<size_t>(b) * dim + d);
    }
    Real v = static_cast<Real>(x) / denom;
    out[e] = apply_tent ? tent(v) : v;
  }
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float fmod1(float x) { return fmodf(x, 1.0f); }
__device__ __forceinline__ double fmod1(double x) { return fmod(x, 1.0); }

// tf.math.floormod(x, 1): C fmod, moved into [0, 1) when negative.
template <typename Real>
__device__ __forceinline__ Real floormod1(Real x) {
  Real r = fmod1(x);
  return (r != Real(0) && r < Real(0)) ? add_rn(r, Real(1)) : r;
}

// lattice_rule.py:203-229: floormod(real(idx) * floormod(z / n, 1) + shift, 1);
// the products and sums are separately rounded (no FMA contraction), as TF's are.
template <typename Int, typename Real>
__global__ void lattice_rule_kernel(const Real* __restrict__ scaled, const Real* __restrict__ shift,
                                    const int64_t* __restrict__ seq, uint64_t first_index,
                                    uint64_t count, int dim, int apply_tent,
                                    Real* __restrict__ out) {
  const uint64_t total = count * static_cast<uint64_t>(dim);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += stride) {
    const uint64_t i = e / dim;
    const int d = static_cast<int>(e - i * dim);
    const Int idx = static_cast<Int>(seq ? seq[i] : static_cast<int64_t>(first_index + i));
    Real p = mul_rn(static_cast<Real>(idx), __ldg(scaled + d));
    if (shift) p = add_rn(p, __ldg(shift + d));
    p = floormod1(p);
    out[e] = apply_tent ? tent(p) : p;
  }
}

template <typename T>
int upload(const std::vector<T>& host, T** dev, cudaStream_t s) {
  TQF_CUDA_OK(cudaMalloc(dev, host.size() * sizeof(T)));
  cudaError_t e = cudaMemcpyAsync(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) {
    cudaFree(*dev);
    *dev = nullptr;
    return cuda_fail(e, "cudaMemcpyAsync(qmc table)");
  }
  return TQF_OK;
}

template <typename Int, typename Real>
int launch_digital_net(const int64_t* g, const int64_t* shift, int dim, int m, const int64_t* seq,
                       uint64_t first_index, uint64_t count, int num_digits, int apply_tent,
                       Real* out, cudaStream_t s) {
  std::vector<Int> gt(static_cast<size_t>(m > 0 ? m : 1) * dim, 0), sh(dim, 0);
  for (int d = 0; d < dim; ++d) {
    for (int b = 0; b < m; ++b) gt[static_cast<size_t>(b) * dim + d] = static_cast<Int>(g[static_cast<size_t>(d) * m + b]);
    if (shift) sh[d] = static_cast<Int>(shift[d]);
  }
  // tf.bitwise.left_shift(1, num_digits) in int_dtype, then cast to real_dtype
  // (digital_net.py:411-417): wraps to the sign bit at num_digits = bits - 1.
  using UInt = typename std::make_unsigned<Int>::type;
  const Int max_binary_point = static_cast<Int>(static_cast<UInt>(1) << num_digits);
  Int *gt_dev = nullptr, *sh_dev = nullptr;
  int rc = upload(gt, &gt_dev, s);
  if (rc != TQF_OK) return rc;
  rc = upload(sh, &sh_dev, s);
  if (rc != TQF_OK) {
    cudaFree(gt_dev);
    return rc;
  }
  digital_net_kernel<Int, Real><<<qmc_grid(count * dim, 256), 256, 0, s>>>(
      gt_dev, sh_dev, seq, first_index, count, dim, m, static_cast<Real>(max_binary_point),
      apply_tent, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the host tables go out of scope
  cudaFree(gt_dev);
  cudaFree(sh_dev);
  if (e != cudaSuccess) return cuda_fail(e, "digital_net_kernel");
  return TQF_OK;
}

template <typename Int, typename Real>
int launch_lattice(const int64_t* gv, int dim, int64_t num_results, const double* shift,
                   const int64_t* seq, uint64_t first_index, uint64_t count, int apply_tent,
                   Real* out, cudaStream_t s) {
  std::vector<Real> scaled(dim), sh(dim);
  const Real n = static_cast<Real>(static_cast<Int>(num_results));
  for (int d = 0; d < dim; ++d) {
    // tf.divide(cast(z), cast(n)) then floormod(., 1)  (lattice_rule.py:206-212)
    volatile Real q = static_cast<Real>(static_cast<Int>(gv[d])) / n;
    Real r = std::fmod(static_cast<Real>(q), static_cast<Real>(1));
    if (r != 0 && r < 0) r += 1;
    scaled[d] = r;
    if (shift) sh[d] = static_cast<Real>(shift[d]);
  }
  Real *scaled_dev = nullptr, *sh_dev = nullptr;
  int rc = upload(scaled, &scaled_dev, s);
  if (rc != TQF_OK) return rc;
  if (shift) {
    rc = upload(sh, &sh_dev, s);
    if (rc != TQF_OK) {
      cudaFree(scaled_dev);
      return rc;
    }
  }
  lattice_rule_kernel<Int, Real><<<qmc_grid(count * dim, 256), 256, 0, s>>>(
      scaled_dev, sh_dev, seq, first_index, count, dim, apply_tent, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(scaled_dev);
  if (sh_dev) cudaFree(sh_dev);
  if (e != cudaSuccess) return cuda_fail(e, "lattice_rule_kernel");
  return TQF_OK;
}

}  // namespace
}  // namespace tqf

using namespace tqf;  // NOLINT

extern "C" {

int tqf_philox_uniform_int_fill(const uint32_t key[2], const uint32_t counter[4], int64_t minval,
                                int64_t maxval, uint64_t num_elements, int int_bits, void* out_dev,
                                void* stream) {
  TQF_NVTX("tqf_philox_uniform_int_fill");
  TQF_REQUIRE(key && counter, "null key/counter");
  TQF_REQUIRE(int_bits == 32 || int_bits == 64, "int_bits must be 32 or 64");
  TQF_REQUIRE(maxval > minval, "maxval must exceed minval");
  if (num_elements == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  const PhiloxKey k{key[0], key[1]};
  const PhiloxCtr c{counter[0], counter[1], counter[2], counter[3]};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int_bits == 32) {
    TQF_REQUIRE(minval >= INT32_MIN && maxval <= INT32_MAX, "int32 bounds out of range");
    const uint32_t range = static_cast<uint32_t>(static_cast<int32_t>(maxval)) -
                           static_cast<uint32_t>(static_cast<int32_t>(minval));
    philox_uniform_int32_kernel<<<qmc_grid(num_elements / 4 + 1, 256), 256, 0, s>>>(
        k, c, static_cast<int32_t>(minval), range, num_elements, static_cast<int32_t*>(out_dev));
  } else {
    const uint64_t range = static_cast<uint64_t>(maxval) - static_cast<uint64_t>(minval);
    philox_uniform_int64_kernel<<<qmc_grid(num_elements / 2 + 1, 256), 256, 0, s>>>(
        k, c, minval, range, num_elements, static_cast<int64_t*>(out_dev));
  }
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_qmc_sobol_generating_matrices(const uint32_t* poly_a, const uint8_t* degree,
                                      const uint32_t* m_init, int num_rows, int dim,
                                      int log_num_results, int num_digits, int64_t* out) {
  TQF_REQUIRE(poly_a && degree && m_init && out, "null table");
  TQF_REQUIRE(dim >= 1 && dim - 1 <= num_rows, "dim out of range of the Joe-Kuo table");
  TQF_REQUIRE(log_num_results >= 0 && log_num_results < 32, "log2(num_results) must be less than 32");
  TQF_REQUIRE(num_digits >= log_num_results && num_digits < 63,
              "num_digits must be in [log2(num_results), 63)");
  const int m = log_num_results;
  // first coordinate: the identity (sobol.py:221-243)
  for (int j = 0; j < m; ++j) out[j] = int64_t{1} << (num_digits - 1 - j);
  for (int k = 0; k + 1 < dim; ++k) {
    int64_t* row = out + static_cast<size_t>(k + 1) * m;
    const int deg = degree[k];
    const uint32_t poly = (1u << deg) + 2u * poly_a[k] + 1u;           // sobol_impl.py:257
    // initial_matrices (sobol.py:318-321): the 18 tabulated m_i, zero beyond
    for (int j = 0; j < m; ++j) {
      const int64_t mi = j < 18 ? m_init[static_cast<size_t>(k) * 18 + j] : 0;
      row[j] = mi << (num_digits - 1 - j);
    }
    // the while loop of sobol.py:353-394, column by column
    for (int column = 0; column + 1 < m; ++column) {
      const int64_t cv = row[column];
      const int lo = (deg > column + 1) ? deg : column + 1;
      for (int i = lo; i <= column + deg && i < m; ++i) {
        const int64_t base = (i == column + deg) ? (cv >> deg) : row[i];
        const int bit = column + deg - i;
        row[i] = base ^ (((poly >> bit) & 1u) ? cv : 0);
      }
    }
  }
  return TQF_OK;
}

int tqf_qmc_scramble_generating_matrices(const int64_t* generating_matrices,
                                         const int64_t* scrambling_matrices, int dim,
                                         int num_columns, int scrambling_columns, int num_digits,
                                         int64_t* out) {
  TQF_REQUIRE(generating_matrices && scrambling_matrices && out, "null matrices");
  TQF_REQUIRE(dim >= 1 && num_columns >= 0, "bad shape");
  TQF_REQUIRE(num_digits >= 1 && num_digits <= scrambling_columns && num_digits < 64,
              "num_digits must be in [1, number of scrambling columns]");
  // digital_net.py:497-525: out = XOR_shift filter(S[:, shift] >> shift, G, num_digits-1-shift)
  for (int d = 0; d < dim; ++d) {
    for (int c = 0; c < num_columns; ++c) {
      const int64_t g = generating_matrices[static_cast<size_t>(d) * num_columns + c];
      int64_t acc = 0;
      for (int shift = 0; shift < num_digits; ++shift) {
        if ((g >> (num_digits - 1 - shift)) & 1) {
          acc ^= scrambling_matrices[static_cast<size_t>(d) * scrambling_columns + shift] >> shift;
        }
      }
      out[static_cast<size_t>(d) * num_columns + c] = acc;
    }
  }
  return TQF_OK;
}

int tqf_qmc_digital_net_fill(const int64_t* generating_matrices, int dim, int num_columns,
                             int log_num_results, const int64_t* digital_shift,
                             const int64_t* sequence_indices_dev, uint64_t first_index,
                             uint64_t count, int num_digits, int int_bits, int apply_tent_transform,
                             int dtype, void* out_dev, void* stream) {
  TQF_NVTX("tqf_qmc_digital_net_fill");
  TQF_REQUIRE((generating_matrices || num_columns == 0) && dim >= 1,
              "null generating matrices / bad dim");
  TQF_REQUIRE(int_bits == 32 || int_bits == 64, "int_bits must be 32 or 64");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  TQF_REQUIRE(log_num_results >= 0 && log_num_results < 32, "log2(num_results) must be less than 32");
  TQF_REQUIRE(log_num_results <= num_columns, "generating matrices have too few columns");
  TQF_REQUIRE(num_digits >= 1 && num_digits < int_bits, "num_digits out of range of the integer type");
  if (count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // only the first log_num_results columns take part (digital_net.py:383-409)
  std::vector<int64_t> g(static_cast<size_t>(dim) * (log_num_results > 0 ? log_num_results : 1), 0);
  for (int d = 0; d < dim; ++d)
    for (int b = 0; b < log_num_results; ++b)
      g[static_cast<size_t>(d) * log_num_results + b] =
          generating_matrices[static_cast<size_t>(d) * num_columns + b];
  const int m = log_num_results;
  if (int_bits == 32 && dtype == TQF_F32)
    return launch_digital_net<int32_t, float>(g.data(), digital_shift, dim, m, sequence_indices_dev,
                                              first_index, count, num_digits, apply_tent_transform,
                                              static_cast<float*>(out_dev), s);
  if (int_bits == 32)
    return launch_digital_net<int32_t, double>(g.data(), digital_shift, dim, m, sequence_indices_dev,
                                               first_index, count, num_digits, apply_tent_transform,
                                               static_cast<double*>(out_dev), s);
  if (dtype == TQF_F32)
    return launch_digital_net<int64_t, float>(g.data(), digital_shift, dim, m, sequence_indices_dev,
                                              first_index, count, num_digits, apply_tent_transform,
                                              static_cast<float*>(out_dev), s);
  return launch_digital_net<int64_t, double>(g.data(), digital_shift, dim, m, sequence_indices_dev,
                                             first_index, count, num_digits, apply_tent_transform,
                                             static_cast<double*>(out_dev), s);
}

int tqf_qmc_lattice_rule_fill(const int64_t* generating_vectors, int dim, int64_t num_results,
                              const double* additive_shift, const int64_t* sequence_indices_dev,
                              uint64_t first_index, uint64_t count, int int_bits,
                              int apply_tent_transform, int dtype, void* out_dev, void* stream) {
  TQF_NVTX("tqf_qmc_lattice_rule_fill");
  TQF_REQUIRE(generating_vectors && dim >= 1, "null generating vectors / bad dim");
  TQF_REQUIRE(int_bits == 32 || int_bits == 64, "int_bits must be 32 or 64");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  TQF_REQUIRE(num_results > 0, "num_results must be positive");
  if (count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int_bits == 32 && dtype == TQF_F32)
    return launch_lattice<int32_t, float>(generating_vectors, dim, num_results, additive_shift,
                                          sequence_indices_dev, first_index, count,
                                          apply_tent_transform, static_cast<float*>(out_dev), s);
  if (int_bits == 32)
    return launch_lattice<int32_t, double>(generating_vectors, dim, num_results, additive_shift,
                                           sequence_indices_dev, first_index, count,
                                           apply_tent_transform, static_cast<double*>(out_dev), s);
  if (dtype == TQF_F32)
    return launch_lattice<int64_t, float>(generating_vectors, dim, num_results, additive_shift,
                                          sequence_indices_dev, first_index, count,
                                          apply_tent_transform, static_cast<float*>(out_dev), s);
  return launch_lattice<int64_t, double>(generating_vectors, dim, num_results, additive_shift,
                                         sequence_indices_dev, first_index, count,
                                         apply_tent_transform, static_cast<double*>(out_dev), s);
}

}  // extern "C"
