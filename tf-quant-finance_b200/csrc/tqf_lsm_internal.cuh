// Shared internals of the Longstaff-Schwartz passes (tqf_lsm.cu: one launch per
// exercise date; tqf_lsm_persist.cu: the whole backward induction in one
// cooperative launch): constants, the K x K regression solve and the handle.
#pragma once

#include <cstdint>
#include <vector>

#include "tqf_common.cuh"
#include "tqf_peer.cuh"

namespace tqf {

constexpr int kLsmBlock = 256;
constexpr int kLsmFastK = 6;                                  // register path
constexpr int kLsmFastNS = kLsmFastK * (kLsmFastK + 1) / 2 + kLsmFastK;  // 27
constexpr int kLsmMaxDim = 8;
constexpr int kLsmMaxK = 128;
constexpr int kLsmTile = 32;   // paths per tile of the generic path

// beta = pinv(X'X) X'y for the packed layout (K <= 6): cyclic Jacobi
// eigen-decomposition of the symmetric PSD matrix, fully unrolled so that the
// matrices live in registers; eigenvalues below rcond * max eigenvalue are
// dropped, as tf.linalg.pinv / numpy.linalg.pinv do with singular values
// (lsm.py:369-377).  One thread per payoff.
template <int K>
__device__ void lsm_solve_one(const double* __restrict__ sp, double rcond, int round_to_float,
                              double* __restrict__ beta) {
  double a[K][K], v[K][K], rhs[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
#pragma unroll
    for (int j = i; j < K; ++j) {
      // packed index of (i, j) in the 6 x 6 upper triangle
      double x = sp[i * kLsmFastK - i * (i - 1) / 2 + (j - i)];
      if (round_to_float) x = static_cast<double>(static_cast<float>(x));
      a[i][j] = x;
      a[j][i] = x;
    }
    double r = sp[kLsmFastK * (kLsmFastK + 1) / 2 + i];
    if (round_to_float) r = static_cast<double>(static_cast<float>(r));
    rhs[i] = r;
#pragma unroll
    for (int j = 0; j < K; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  }
  // Fast path: Cholesky.  With A = L L', trace(A^-1) = ||L^-1||_F^2 >= 1/lambda_min
  // and trace(A) >= lambda_max, so 1/trace(A^-1) > rcond trace(A) proves that no
  // singular value falls under the pinv cut-off, i.e. pinv(A) = A^-1 exactly and
  // beta = L^-T L^-1 b.  (Tight within a factor K^2; everything else -- rank
  // deficient or borderline -- takes the eigen-decomposition below.)
  {
    double l[K][K], li[K][K];
    bool ok = true;
    double tr = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      tr += a[j][j];
      double d = a[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) d = fma(-l[j][k], l[j][k], d);
      ok = ok && (d > 0.0);
      const double dj = d > 0.0 ? d : 1.0;
      const double inv = rsqrt(dj);
      l[j][j] = dj * inv;
      li[j][j] = inv;
#pragma unroll
      for (int i = j + 1; i < K; ++i) {
        double v2 = a[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) v2 = fma(-l[i][k], l[j][k], v2);
        l[i][j] = v2 * inv;
      }
    }
    // L^-1 (lower triangular)
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
      for (int i = j + 1; i < K; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = j; k < i; ++k) acc = fma(l[i][k], li[k][j], acc);
        li[i][j] = -acc * li[i][i];
      }
    }
    double tr_inv = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) tr_inv = fma(li[i][j], li[i][j], tr_inv);
    if (ok && tr_inv > 0.0 && 1.0 / tr_inv > rcond * tr) {
      double y[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) acc = fma(li[i][j], rhs[j], acc);
        y[i] = acc;
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = j; i < K; ++i) acc = fma(li[i][j], y[i], acc);
        if (round_to_float) acc = static_cast<double>(static_cast<float>(acc));
        beta[j] = acc;
      }
      return;
    }
  }
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      diag += a[i][i] * a[i][i];
#pragma unroll
      for (int j = i + 1; j < K; ++j) off += a[i][j] * a[i][j];
    }
    if (off <= 1e-300 || off <= 1e-32 * diag) break;
#pragma unroll
    for (int p = 0; p < K - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < K; ++q) {
        const double apq = a[p][q];
        if (apq != 0.0) {
          const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = rsqrt(t * t + 1.0), sn = t * c;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const double akp = a[k][p], akq = a[k][q];
            a[k][p] = c * akp - sn * akq;
            a[k][q] = sn * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const double apk = a[p][k], aqk = a[q][k];
            a[p][k] = c * apk - sn * aqk;
            a[q][k] = sn * apk + c * aqk;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const double vkp = v[k][p], vkq = v[k][q];
            v[k][p] = c * vkp - sn * vkq;
            v[k][q] = sn * vkp + c * vkq;
          }
        }
      }
    }
  }
  double lmax = 0.0;
#pragma unroll
  for (int i = 0; i < K; ++i) lmax = fmax(lmax, fabs(a[i][i]));
  const double cutoff = rcond * lmax;
  double coef[K];
#pragma unroll
  for (int e = 0; e < K; ++e) {
    double proj = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) proj += v[j][e] * rhs[j];
    const double lam = a[e][e];
    coef[e] = fabs(lam) > cutoff ? proj / lam : 0.0;
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < K; ++e) s += v[i][e] * coef[e];
    if (round_to_float) s = static_cast<double>(static_cast<float>(s));
    beta[i] = s;
  }
}

}  // namespace tqf

struct tqf_lsm {
  tqf_lsm_desc desc;
  int K, NS, grid, grid_aux;
  bool fast;
  void* w_dev;           // Real [B][N]
  int* exponents_dev;    // [K][dim]
  double* strikes_dev;   // [B]
  double* partials_dev;  // [grid][B][max(NS, T*dim, 2)]
  size_t partials_doubles;
  int* times_dev;
  int times_cap;
  bool external_w, external_partials;
  bool tabulated;                   // exercise values and / or per-path ratios given
  // fused solve (tqf_lsm_set_fused_solve)
  unsigned int* ticket_dev;
  double* fused_sums_dev;
  double* fused_beta_dev;
  double fused_rcond;
  bool last_step_solved;
  // peer exchange (tqf_lsm_set_peer_exchange)
  int peer_rank, peer_world;
  unsigned long long peer_epoch;
  unsigned char* peer_bufs[tqf::kLsmMaxPeers];
  std::vector<int>* exercise_times; // slot -> time index (tabulated mode)
  // persistent backward induction (tqf_lsm_persist.cu): {arrival counter, release
  // flag, status} of its grid barrier, zeroed before every run
  unsigned long long* ctrl_dev;
  double strike0;                   // strike of payoff 0 (host copy)
};


namespace tqf {

// The vectorised single-asset kernel (the one with the fused solve) applies.
inline bool lsm_vec_ok(const tqf_lsm* h) {
  const tqf_lsm_desc& d = h->desc;
  const size_t esz = d.dtype == TQF_F64 ? 8 : 4;
  return !h->tabulated && h->fast && d.dim == 1 && d.stride_path == 1 && (d.num_paths % 2) == 0 &&
         d.num_paths > 0 && d.num_paths < (1ull << 32) &&
         (reinterpret_cast<uintptr_t>(d.paths_dev) % (2 * esz)) == 0 &&
         (reinterpret_cast<uintptr_t>(h->w_dev) % (2 * esz)) == 0 && (d.stride_time % 2) == 0 &&
         (d.stride_batch % 2) == 0;
}


// tqf_lsm_persist.cu
int lsm_run_persistent(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                       const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                       double rcond, uint64_t skip_below, double* value_sums_dev,
                       double* beta_dev, double* history_dev, cudaStream_t stream);
bool lsm_persistent_ok(const tqf_lsm* h);

}  // namespace tqf
