// Hull-White discount curves along simulated short-rate paths.
//
// Replaces the device work of `sample_discount_curve_paths` /
// `_bond_reconstitution` (models/hull_white/vector_hull_white.py:451-592,
// 783-814): the reference broadcasts eight elementwise TensorFlow ops over the
// [num_samples, num_curve_times, num_sim_times, dim] grid,
//   P(t_j, t_j + tau_i) = P0(t_j + tau_i) / P0(t_j) exp(-x G - y G^2 / 2),
//   x = r(t_j) - f(0, t_j),  G = (1 - e^{-a tau_i}) / a,
// each materialising a tensor of that size.  Everything that does not depend on
// the path is folded on the host into two [m, k, dim] tables
//   A = P0(t + tau) / P0(t) exp(-y G^2 / 2)   and   G,
// and ONE kernel writes  P = A exp(-(r - f0) G): one exponential and 8 (4) bytes
// stored per output element, nothing else touches HBM except the rates (1 / m of
// the output).  The rates arrive time-major (the layout the path engine writes:
// consecutive paths are adjacent), the output is path-major like the
// reference's tensor: a 32 x 32 shared-memory tile transposes between the two so
// that both the loads and the stores are coalesced.
#include "tqf_common.cuh"

namespace tqf {

constexpr int kHwTile = 32;

template <typename Real>
__global__ void __launch_bounds__(kHwTile * 8)
hw_discount_curves_kernel(const Real* __restrict__ rates, int64_t rs_path, int64_t rs_time,
                          int64_t rs_dim, const double* __restrict__ f0,
                          const double* __restrict__ coef_a, const double* __restrict__ coef_g,
                          uint64_t num_paths, int m, int k, int dim, Real* __restrict__ out) {
  __shared__ double s_x[kHwTile][kHwTile + 1];          // [column][path]
  const int kd = k * dim;
  const uint64_t n0 = static_cast<uint64_t>(blockIdx.x) * kHwTile;
  const int c0 = blockIdx.y * kHwTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // x = r - f(0, t): coalesced along the path axis
  for (int r = ty; r < kHwTile; r += 8) {
    const int c = c0 + r;
    const uint64_t n = n0 + tx;
    double x = 0.0;
    if (c < kd && n < num_paths) {
      const int j = c / dim, d = c - j * dim;
      x = static_cast<double>(rates[static_cast<int64_t>(n) * rs_path + j * rs_time + d * rs_dim]) - f0[c];
    }
    s_x[r][tx] = x;
  }
  __syncthreads();
  const int c = c0 + tx;
  if (c >= kd) return;
  for (int i = 0; i < m; ++i) {
    const double a = coef_a[static_cast<size_t>(i) * kd + c];
    const double g = coef_g[static_cast<size_t>(i) * kd + c];
    for (int r = ty; r < kHwTile; r += 8) {
      const uint64_t n = n0 + r;
      if (n < num_paths)
        out[(n * m + i) * static_cast<uint64_t>(kd) + c] = static_cast<Real>(a * exp(-s_x[tx][r] * g));
    }
  }
}

// HJM discount curves (`_bond_reconstitution`, hjm/quasi_gaussian_hjm.py:499-525)
// along simulated factor paths: the factors act TOGETHER inside one exponential,
//   P(t_j, t_j + tau_i) = A_ij exp(-sum_d G_ijd x_d(t_j)),
//   A = P0(t + tau) / P0(t) exp(-G' y(t) G / 2)  (y deterministic: folded on the host).
// Same tile transpose as above; out is [num_paths][m][k].
template <typename Real>
__global__ void __launch_bounds__(kHwTile * 8)
hjm_discount_curves_kernel(const Real* __restrict__ x, int64_t xs_path, int64_t xs_time,
                           int64_t xs_dim, const double* __restrict__ coef_a,
                           const double* __restrict__ coef_g, uint64_t num_paths, int m, int k,
                           int nf, Real* __restrict__ out) {
  __shared__ double s_x[3][kHwTile][kHwTile + 1];       // [factor][time column][path]
  const uint64_t n0 = static_cast<uint64_t>(blockIdx.x) * kHwTile;
  const int c0 = blockIdx.y * kHwTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < kHwTile; r += 8) {
    const int j = c0 + r;
    const uint64_t n = n0 + tx;
    for (int d = 0; d < nf; ++d) {
      double v = 0.0;
      if (j < k && n < num_paths)
        v = static_cast<double>(x[static_cast<int64_t>(n) * xs_path + j * xs_time + d * xs_dim]);
      s_x[d][r][tx] = v;
    }
  }
  __syncthreads();
  const int j = c0 + tx;
  if (j >= k) return;
  for (int i = 0; i < m; ++i) {
    const double a = coef_a[static_cast<size_t>(i) * k + j];
    double g[3];
    for (int d = 0; d < nf; ++d) g[d] = coef_g[(static_cast<size_t>(i) * k + j) * nf + d];
    for (int r = ty; r < kHwTile; r += 8) {
      const uint64_t n = n0 + r;
      if (n < num_paths) {
        double e = 0.0;
        for (int d = 0; d < nf; ++d) e = fma(-s_x[d][tx][r], g[d], e);
        out[(n * m + i) * static_cast<uint64_t>(k) + j] = static_cast<Real>(a * exp(e));
      }
    }
  }
}

// Exercise values of Bermudan swaptions on Hull-White paths: the tabulated
// payoff that hull_white/swaption.py:608-724 builds with a gather of the
// [N, m, k] bond tensor, a weighted sum over the payments and a scatter to the
// simulation times (`_map_payoff_to_sim_times`):
//   values[u][n][b] = relu(1 - sum_j coef[b][e][j] exp(kk[b][e][j] - g[b][e][j] x[n][u])),
// u = ex_slot[b][e].  One thread per path; entries of dates a swaption cannot be
// exercised on stay untouched (the caller zero-fills), later exercise dates of
// one swaption that map to the same slot overwrite earlier ones.
template <typename Real>
__global__ void __launch_bounds__(256)
hw_exercise_values_kernel(const Real* __restrict__ x, int64_t xs_path, int64_t xs_slot,
                          const double* __restrict__ tab_g, const double* __restrict__ tab_k,
                          const double* __restrict__ tab_c, const int* __restrict__ ex_slot,
                          uint64_t num_paths, int nb, int n_ex, int m, Real* __restrict__ values) {
  const uint64_t n = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= num_paths) return;
  for (int b = 0; b < nb; ++b)
    for (int e = 0; e < n_ex; ++e) {
      const int u = ex_slot[b * n_ex + e];
      const double xv = static_cast<double>(x[static_cast<int64_t>(n) * xs_path + u * xs_slot]);
      const size_t row = (static_cast<size_t>(b) * n_ex + e) * m;
      double acc = 0.0;
      for (int j = 0; j < m; ++j)
        acc = fma(tab_c[row + j], exp(fma(-tab_g[row + j], xv, tab_k[row + j])), acc);
      const double swap = 1.0 - acc;
      values[(static_cast<uint64_t>(u) * num_paths + n) * nb + b] =
          static_cast<Real>(swap > 0.0 ? swap : 0.0);
    }
}

}  // namespace tqf

using namespace tqf;

extern "C" int tqf_hw_exercise_values(const void* x_dev, int64_t xs_path, int64_t xs_slot,
                                      const double* g_dev, const double* k_dev,
                                      const double* coef_dev, const int32_t* ex_slot_dev,
                                      uint64_t num_paths, int num_swaptions, int num_exercise,
                                      int num_payments, int dtype, void* values_dev,
                                      void* stream) {
  TQF_REQUIRE(x_dev && g_dev && k_dev && coef_dev && ex_slot_dev && values_dev, "null argument");
  TQF_REQUIRE(num_swaptions >= 1 && num_exercise >= 1 && num_payments >= 1, "empty axis");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "bad dtype");
  if (num_paths == 0) return TQF_OK;
  const uint64_t blocks = (num_paths + 255) / 256;
  TQF_REQUIRE(blocks < (1ull << 31), "too many paths for one launch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == TQF_F64)
    hw_exercise_values_kernel<double><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        static_cast<const double*>(x_dev), xs_path, xs_slot, g_dev, k_dev, coef_dev, ex_slot_dev,
        num_paths, num_swaptions, num_exercise, num_payments, static_cast<double*>(values_dev));
  else
    hw_exercise_values_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        static_cast<const float*>(x_dev), xs_path, xs_slot, g_dev, k_dev, coef_dev, ex_slot_dev,
        num_paths, num_swaptions, num_exercise, num_payments, static_cast<float*>(values_dev));
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

extern "C" int tqf_hw_discount_curves(const void* rates_dev, int64_t rs_path, int64_t rs_time,
                                      int64_t rs_dim, const double* f0_dev,
                                      const double* coef_a_dev, const double* coef_g_dev,
                                      uint64_t num_paths, int m, int k, int dim, int dtype,
                                      void* out_dev, void* stream) {
  TQF_REQUIRE(rates_dev && f0_dev && coef_a_dev && coef_g_dev && out_dev, "null argument");
  TQF_REQUIRE(m >= 1 && k >= 1 && dim >= 1, "empty curve / time / factor axis");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "bad dtype");
  if (num_paths == 0) return TQF_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libtqf has no CPU fallback");
    return TQF_ERR_CUDA;
  }
  const uint64_t bx = (num_paths + kHwTile - 1) / kHwTile;
  TQF_REQUIRE(bx < (1ull << 31), "too many paths for one launch");
  const dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>((k * dim + kHwTile - 1) / kHwTile));
  const dim3 block(kHwTile, 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == TQF_F64)
    hw_discount_curves_kernel<double><<<grid, block, 0, s>>>(
        static_cast<const double*>(rates_dev), rs_path, rs_time, rs_dim, f0_dev, coef_a_dev,
        coef_g_dev, num_paths, m, k, dim, static_cast<double*>(out_dev));
  else
    hw_discount_curves_kernel<float><<<grid, block, 0, s>>>(
        static_cast<const float*>(rates_dev), rs_path, rs_time, rs_dim, f0_dev, coef_a_dev,
        coef_g_dev, num_paths, m, k, dim, static_cast<float*>(out_dev));
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

extern "C" int tqf_hjm_discount_curves(const void* x_dev, int64_t xs_path, int64_t xs_time,
                                       int64_t xs_dim, const double* coef_a_dev,
                                       const double* coef_g_dev, uint64_t num_paths, int m,
                                       int k, int num_factors, int dtype, void* out_dev,
                                       void* stream) {
  TQF_NVTX("tqf_hjm_discount_curves");
  TQF_REQUIRE(x_dev && coef_a_dev && coef_g_dev && out_dev, "null argument");
  TQF_REQUIRE(m >= 1 && k >= 1 && num_factors >= 1 && num_factors <= 3,
              "empty curve / time axis or more than 3 factors");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "bad dtype");
  if (num_paths == 0) return TQF_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libtqf has no CPU fallback");
    return TQF_ERR_CUDA;
  }
  const uint64_t bx = (num_paths + kHwTile - 1) / kHwTile;
  TQF_REQUIRE(bx < (1ull << 31), "too many paths for one launch");
  const dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>((k + kHwTile - 1) / kHwTile));
  const dim3 block(kHwTile, 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == TQF_F64)
    hjm_discount_curves_kernel<double><<<grid, block, 0, s>>>(
        static_cast<const double*>(x_dev), xs_path, xs_time, xs_dim, coef_a_dev, coef_g_dev,
        num_paths, m, k, num_factors, static_cast<double*>(out_dev));
  else
    hjm_discount_curves_kernel<float><<<grid, block, 0, s>>>(
        static_cast<const float*>(x_dev), xs_path, xs_time, xs_dim, coef_a_dev, coef_g_dev,
        num_paths, m, k, num_factors, static_cast<float*>(out_dev));
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}
