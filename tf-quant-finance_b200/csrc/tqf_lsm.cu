// Longstaff-Schwartz regression passes on materialised paths.
//
// Replaces the device work of models/longstaff_schwartz/lsm.py:231-436 of the
// reference (payoff_fn, basis_fn, the masked X'X / X'y matmuls and the
// tf.where updates, each a full pass over [N] tensors there).  Per exercise
// date ONE streaming pass updates the merged state W = cashflow + values
//   W' = exercise_value > relu(X beta) ? exercise_value : ratio W
// (exact because one of cashflow / values is always zero, lsm.py:391-399) and
// accumulates the normal equations of the NEXT (earlier) date from W'.  The
// K x K pseudo-inverse happens between passes on the reduced sums (and, on
// several GPUs, after the all-reduce of those sums).
//
// HBM-bound: per path and date it reads x at two time columns and reads +
// writes W (4 x 8 B in fp64).  Paths are expected time-major (stride_path = 1),
// the layout tqf_plan_paths writes; any strides work.
#include <cstring>
#include <new>
#include <vector>

#include "tqf_lsm_internal.cuh"

namespace tqf {

template <typename Real>
struct LsmArgs {
  const Real* paths;
  int64_t stride_path, stride_time, stride_dim, stride_batch;
  Real* w;              // [B][N]
  uint64_t num_paths;   // local
  uint64_t path_offset; // global index of local path 0
  uint64_t num_calib;   // regression uses global paths < num_calib
  int dim, K, batch;
  const int* exponents;    // device [K][dim]
  const double* strikes;   // device [B]
  // update part (all pointers device memory)
  int do_update, t_update;
  const double* mean_update;  // [B][mean_stride] -> first `dim` entries used
  const double* beta;         // [B][K]
  const double* ratio_update; // [B]
  // accumulate part
  int do_acc, t_acc;
  const double* mean_acc;
  const double* ratio_acc;    // [B]
  int64_t mean_stride;        // doubles between the means of consecutive payoffs
  double* partials;           // device [gridDim.x][B][NS]
  int NS;
  // tabulated exercise values [T][B][N] / per-path ratios [T][N] (either may be null)
  const Real* ev_tab;
  const Real* ratio_path;
  int slot_update, slot_acc;  // exercise-date slots of t_update / t_acc
  // fused regression solve (vectorised single-asset kernel): the last CTA to
  // finish reduces the partials into `sums_out` and writes beta for `t_acc`
  unsigned int* ticket;       // device counter, zero between launches (null: not fused)
  double* sums_out;           // [B][kLsmFastNS]
  double* beta_out;           // [B][K]
  double rcond;
  int round_to_float;
  // exchange of the reduced sums between the GPUs of one box through peer
  // memory (NVLink), inside the same tail -- no NCCL call per exercise date
  int peer_rank, peer_world;            // peer_world <= 1: single GPU
  unsigned long long peer_epoch;        // strictly increasing launch number, equal on all ranks
  unsigned char* peer_bufs[kLsmMaxPeers];  // exchange buffer of every rank, mapped here
};

template <typename Real>
__device__ __forceinline__ Real lsm_payoff(const LsmArgs<Real>& A, const Real* xp, int b) {
  // relu(strike - mean_j x_j) (payoff_utils.py:95-97)
  Real s = 0;
  for (int j = 0; j < A.dim; ++j) s += xp[j * A.stride_dim];
  const Real avg = s / static_cast<Real>(A.dim);
  const Real v = static_cast<Real>(A.strikes[b]) - avg;
  return v > Real(0) ? v : Real(0);
}

// phi[k] = prod_j (x_j - mean_j)^e[k][j]  (lsm.py:110-124)
template <typename Real>
__device__ __forceinline__ void lsm_basis(const LsmArgs<Real>& A, const Real* xp,
                                          const double* mean, Real* phi) {
  if (A.dim == 1) {
    const Real c = xp[0] - static_cast<Real>(mean[0]);
    // exponents of the 1-d basis are 0..K-1 in order
    Real p = 1;
    for (int k = 0; k < A.K; ++k) {
      phi[k] = p;
      p *= c;
    }
    return;
  }
  Real c[kLsmMaxDim];
  for (int j = 0; j < A.dim; ++j) c[j] = xp[j * A.stride_dim] - static_cast<Real>(mean[j]);
  for (int k = 0; k < A.K; ++k) {
    Real p = 1;
    for (int j = 0; j < A.dim; ++j) {
      const int e = A.exponents[k * A.dim + j];
      Real q = 1;
      for (int i = 0; i < e; ++i) q *= c[j];
      p *= q;
    }
    phi[k] = p;
  }
}

// Fast path: K <= 6, accumulators in registers.  grid = (blocks, B).
template <typename Real>
__global__ void __launch_bounds__(kLsmBlock) lsm_step_fast_kernel(const LsmArgs<Real> A) {
  const int b = blockIdx.y;
  const int K = A.K;
  const Real* base = A.paths + b * A.stride_batch;
  Real* w = A.w + static_cast<size_t>(b) * A.num_paths;
  double acc[kLsmFastNS];
#pragma unroll
  for (int i = 0; i < kLsmFastNS; ++i) acc[i] = 0.0;
  double beta[kLsmFastK];
#pragma unroll
  for (int k = 0; k < kLsmFastK; ++k) beta[k] = (A.do_update && k < K) ? A.beta[b * K + k] : 0.0;
  const Real ratio_u = A.do_update ? static_cast<Real>(A.ratio_update[b]) : Real(1);
  const Real ratio_a = A.do_acc ? static_cast<Real>(A.ratio_acc[b]) : Real(1);

  constexpr int U = 4;  // paths in flight per thread
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const double* mean_u = A.mean_update + b * A.mean_stride;
  const double* mean_a = A.mean_acc + b * A.mean_stride;
  for (uint64_t n0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       n0 < A.num_paths; n0 += U * stride) {
    Real wn[U], xu[U], xa[U];
    bool live[U];
    // issue every load of the U paths first (single-asset fast path keeps the
    // two time columns in registers; dim > 1 re-reads inside the helpers)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t n = n0 + u * stride;
      live[u] = n < A.num_paths;
      const Real* xn = base + static_cast<int64_t>(live[u] ? n : 0) * A.stride_path;
      wn[u] = live[u] ? w[n] : Real(0);
      xu[u] = (live[u] && A.do_update) ? xn[A.t_update * A.stride_time] : Real(0);
      xa[u] = (live[u] && A.do_acc) ? xn[A.t_acc * A.stride_time] : Real(0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      const uint64_t n = n0 + u * stride;
      const Real* xn = base + static_cast<int64_t>(n) * A.stride_path;
      Real phi[kLsmFastK];
      if (A.do_update) {
        const Real* xp = xn + A.t_update * A.stride_time;
        Real ev;
        if (A.dim == 1) {
          const Real v = static_cast<Real>(A.strikes[b]) - xu[u];
          ev = v > Real(0) ? v : Real(0);
          if (A.ev_tab)
            ev = A.ev_tab[(static_cast<size_t>(A.slot_update) * A.batch + b) * A.num_paths + n];
          const Real c = xu[u] - static_cast<Real>(mean_u[0]);
          Real pw = 1;
#pragma unroll
          for (int k = 0; k < kLsmFastK; ++k) {
            phi[k] = pw;
            pw *= c;
          }
        } else {
          ev = A.ev_tab
                   ? A.ev_tab[(static_cast<size_t>(A.slot_update) * A.batch + b) * A.num_paths + n]
                   : lsm_payoff(A, xp, b);
          lsm_basis(A, xp, mean_u, phi);
        }
        Real cont = 0;
#pragma unroll
        for (int k = 0; k < kLsmFastK; ++k)
          if (k < K) cont += phi[k] * static_cast<Real>(beta[k]);
        cont = cont > Real(0) ? cont : Real(0);
        const Real ru = A.ratio_path
                            ? A.ratio_path[static_cast<size_t>(A.slot_update + 1) * A.num_paths + n]
                            : ratio_u;
        wn[u] = ev > cont ? ev : ru * wn[u];
        w[n] = wn[u];
      }
      if (A.do_acc) {
        const Real* xp = xn + A.t_acc * A.stride_time;
        Real ev;
        if (A.ev_tab) {
          ev = A.ev_tab[(static_cast<size_t>(A.slot_acc) * A.batch + b) * A.num_paths + n];
        } else if (A.dim == 1) {
          const Real v = static_cast<Real>(A.strikes[b]) - xa[u];
          ev = v > Real(0) ? v : Real(0);
        } else {
          ev = lsm_payoff(A, xp, b);
        }
        const bool use = ev > Real(0) && (A.path_offset + n) < A.num_calib;
        if (use) {
          if (A.dim == 1) {
            const Real c = xa[u] - static_cast<Real>(mean_a[0]);
            Real pw = 1;
#pragma unroll
            for (int k = 0; k < kLsmFastK; ++k) {
              phi[k] = pw;
              pw *= c;
            }
          } else {
            lsm_basis(A, xp, mean_a, phi);
          }
          const Real ra = A.ratio_path
                              ? A.ratio_path[static_cast<size_t>(A.slot_acc + 1) * A.num_paths + n]
                              : ratio_a;
          const double y = static_cast<double>(ra * wn[u]);
          int idx = 0;
#pragma unroll
          for (int i = 0; i < kLsmFastK; ++i) {
#pragma unroll
            for (int j = i; j < kLsmFastK; ++j) {
              if (j < K && i < K)
                acc[idx] += static_cast<double>(phi[i]) * static_cast<double>(phi[j]);
              ++idx;
            }
          }
#pragma unroll
          for (int i = 0; i < kLsmFastK; ++i)
            if (i < K) acc[kLsmFastK * (kLsmFastK + 1) / 2 + i] += static_cast<double>(phi[i]) * y;
        }
      }
    }
  }
  if (A.do_acc) {
    __shared__ double s_red[kLsmBlock / 32][kLsmFastNS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kLsmFastNS; ++i) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLsmFastNS) {
      double v = 0.0;
      for (int wi = 0; wi < kLsmBlock / 32; ++wi) v += s_red[wi][threadIdx.x];
      A.partials[(static_cast<size_t>(blockIdx.x) * A.batch + b) * kLsmFastNS + threadIdx.x] = v;
    }
  }
}

// Single-asset specialisation (the American put of config C5): K compile-time,
// 1 + x + ... + x^(K-1) basis, K (K + 1) / 2 + K accumulators in registers, four
// paths in flight per thread.  Same packed partials layout as the general fast
// kernel (6 x 6 upper triangle + 6), unused entries zero.
template <typename Real, int KT>
__global__ void __launch_bounds__(kLsmBlock, 3) lsm_step_dim1_kernel(const LsmArgs<Real> A) {
  const int b = blockIdx.y;
  const Real* base = A.paths + b * A.stride_batch;
  Real* w = A.w + static_cast<size_t>(b) * A.num_paths;
  constexpr int NA = KT * (KT + 1) / 2 + KT;
  double acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.0;
  Real beta[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) beta[k] = A.do_update ? static_cast<Real>(A.beta[b * KT + k]) : Real(0);
  const Real ratio_u = A.do_update ? static_cast<Real>(A.ratio_update[b]) : Real(1);
  const Real ratio_a = A.do_acc ? static_cast<Real>(A.ratio_acc[b]) : Real(1);
  const Real strike = static_cast<Real>(A.strikes[b]);
  const Real mean_u = A.do_update ? static_cast<Real>(A.mean_update[b * A.mean_stride]) : Real(0);
  const Real mean_a = A.do_acc ? static_cast<Real>(A.mean_acc[b * A.mean_stride]) : Real(0);
  const int64_t off_u = A.t_update * A.stride_time, off_a = A.t_acc * A.stride_time;
  constexpr int U = 4;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t n0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       n0 < A.num_paths; n0 += U * stride) {
    Real wn[U], xu[U], xa[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t n = n0 + u * stride;
      const bool live = n < A.num_paths;
      const Real* xn = base + static_cast<int64_t>(live ? n : 0) * A.stride_path;
      wn[u] = live ? w[n] : Real(0);
      xu[u] = (live && A.do_update) ? xn[off_u] : strike;   // strike -> exercise value 0
      xa[u] = (live && A.do_acc) ? xn[off_a] : strike;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t n = n0 + u * stride;
      if (n >= A.num_paths) continue;
      if (A.do_update) {
        const Real v = strike - xu[u];
        const Real ev = v > Real(0) ? v : Real(0);
        const Real c = xu[u] - mean_u;
        Real cont = beta[KT - 1];
#pragma unroll
        for (int k = KT - 2; k >= 0; --k) cont = fma(cont, c, beta[k]);
        // (Horner; the reference's matmul sums phi_k beta_k -- same value up to rounding)
        cont = cont > Real(0) ? cont : Real(0);
        wn[u] = ev > cont ? ev : ratio_u * wn[u];
        w[n] = wn[u];
      }
      if (A.do_acc) {
        const Real v = strike - xa[u];
        if (v > Real(0) && (A.path_offset + n) < A.num_calib) {
          double phi[KT];
          const double c = static_cast<double>(static_cast<Real>(xa[u] - mean_a));
          phi[0] = 1.0;
#pragma unroll
          for (int k = 1; k < KT; ++k)
            phi[k] = static_cast<double>(static_cast<Real>(static_cast<Real>(phi[k - 1]) * static_cast<Real>(c)));
          const double y = static_cast<double>(ratio_a * wn[u]);
          int idx = 0;
#pragma unroll
          for (int i = 0; i < KT; ++i)
#pragma unroll
            for (int j = i; j < KT; ++j) {
              acc[idx] = fma(phi[i], phi[j], acc[idx]);
              ++idx;
            }
#pragma unroll
          for (int i = 0; i < KT; ++i) acc[KT * (KT + 1) / 2 + i] = fma(phi[i], y, acc[KT * (KT + 1) / 2 + i]);
        }
      }
    }
  }
  if (A.do_acc) {
    __shared__ double s_red[kLsmBlock / 32][NA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLsmFastNS) {
      // map packed (6 x 6) slot -> compact (KT x KT) slot
      const int slot = threadIdx.x;
      int src = -1;
      if (slot < kLsmFastK * (kLsmFastK + 1) / 2) {
        int i = 0, rem = slot;
        while (rem >= kLsmFastK - i) {
          rem -= kLsmFastK - i;
          ++i;
        }
        const int j = i + rem;
        if (i < KT && j < KT) src = i * KT - i * (i - 1) / 2 + (j - i);
      } else {
        const int i = slot - kLsmFastK * (kLsmFastK + 1) / 2;
        if (i < KT) src = KT * (KT + 1) / 2 + i;
      }
      double v = 0.0;
      if (src >= 0)
        for (int wi = 0; wi < kLsmBlock / 32; ++wi) v += s_red[wi][src];
      A.partials[(static_cast<size_t>(blockIdx.x) * A.batch + b) * kLsmFastNS + slot] = v;
    }
  }
}

// Contiguous variant of the single-asset kernel (time-major paths,
// stride_path == 1, even number of paths, 16-byte aligned columns): two paths
// per vector load/store, 32-bit indexing.  The general kernel above spends
// ~135 instructions per path, mostly on 64-bit address arithmetic and
// predication -- enough to be issue-bound below the HBM roofline.
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// Run by the last CTA of a fused pass: fixed-order reduction of the per-CTA
// partial rows, then the K x K solve of every payoff.  Not inlined, so that the
// solver's registers do not weigh on the streaming loop of the caller.
template <int KT>
__device__ __noinline__ void lsm_fused_tail(const PeerK A, const double* partials, int batch,
                                            int num_blocks, double* sums_out, double rcond,
                                            int round_to_float, double* beta_out) {
  // Thread t sums column m = t % 32 (m < M) of the rows r = t / 32, t / 32 + 8, ...
  // with 14 independent L2 loads in flight (the rows were written by other SMs
  // in this launch: read through L2), then the 8 row groups are combined in a
  // fixed order -> reproducible sums.
  const int M = batch * kLsmFastNS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double s_part[kLsmBlock / 32][32];
  constexpr int W = kLsmBlock / 32, U = 14;
  for (int m0 = 0; m0 < M; m0 += 32) {
    const int m = m0 + lane;
    double acc = 0.0;
    if (m < M) {
      int r = warp;
      for (; r + (U - 1) * W < num_blocks; r += U * W) {
        double ld[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ld[u] = __ldcg(partials + static_cast<size_t>(r + u * W) * M + m);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += ld[u];
      }
      for (; r < num_blocks; r += W) acc += __ldcg(partials + static_cast<size_t>(r) * M + m);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && m < M) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < W; ++w) v += s_part[w][lane];
      sums_out[m] = v;
    }
    __syncthreads();
  }
  __syncthreads();
  if (A.peer_world > 1) peer_all_reduce(A, sums_out, M);
  for (int bb = threadIdx.x; bb < batch; bb += kLsmBlock)
    lsm_solve_one<KT>(sums_out + static_cast<size_t>(bb) * kLsmFastNS, rcond, round_to_float,
                      beta_out + static_cast<size_t>(bb) * KT);
}

// L2 residency: the merged state W (8 B per path) is read and written by every
// pass while each path column is read by two consecutive passes only.  W is
// tagged evict_last and the columns evict_first (streaming), so that for sample
// counts whose W fits the 126 MB L2 the passes fetch W from L2, not from HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_keep(const double2* ptr, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_keep(const float2* ptr, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_keep(double2* ptr, double2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;"
               :: "l"(ptr), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(float2* ptr, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;"
               :: "l"(ptr), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

template <typename Real, int KT>
__global__ void __launch_bounds__(kLsmBlock, 3) lsm_step_dim1_vec_kernel(const LsmArgs<Real> A) {
  using V = typename Vec2<Real>::type;
  const uint64_t keep = l2_policy_evict_last();
  const int b = blockIdx.y;
  const Real* base = A.paths + b * A.stride_batch;
  V* __restrict__ w = reinterpret_cast<V*>(A.w + static_cast<size_t>(b) * A.num_paths);
  const V* __restrict__ colu = reinterpret_cast<const V*>(base + A.t_update * A.stride_time);
  const V* __restrict__ cola = reinterpret_cast<const V*>(base + A.t_acc * A.stride_time);
  constexpr int NA = KT * (KT + 1) / 2 + KT;
  double acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.0;
  Real beta[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) beta[k] = A.do_update ? static_cast<Real>(A.beta[b * KT + k]) : Real(0);
  const Real ratio_u = A.do_update ? static_cast<Real>(A.ratio_update[b]) : Real(1);
  const Real ratio_a = A.do_acc ? static_cast<Real>(A.ratio_acc[b]) : Real(1);
  const Real strike = static_cast<Real>(A.strikes[b]);
  const Real mean_u = A.do_update ? static_cast<Real>(A.mean_update[b * A.mean_stride]) : Real(0);
  const Real mean_a = A.do_acc ? static_cast<Real>(A.mean_acc[b * A.mean_stride]) : Real(0);
  const uint32_t npairs = static_cast<uint32_t>(A.num_paths >> 1);
  const uint32_t stride = gridDim.x * blockDim.x;
  const bool calib_all = A.num_calib == ~0ull;
  constexpr int U = 2;  // vector pairs in flight per thread
  for (uint32_t p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < npairs; p0 += U * stride) {
    V wv[U], xu[U], xa[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = p0 + u * stride;
      if (p < npairs) {
        wv[u] = ld_keep(w + p, keep);
        if (A.do_update) xu[u] = __ldcs(colu + p);
        if (A.do_acc) xa[u] = __ldcs(cola + p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = p0 + u * stride;
      if (p >= npairs) continue;
      Real wn[2] = {wv[u].x, wv[u].y};
      const Real xus[2] = {xu[u].x, xu[u].y};
      const Real xas[2] = {xa[u].x, xa[u].y};
      if (A.do_update) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const Real v = strike - xus[e];
          const Real ev = v > Real(0) ? v : Real(0);
          const Real c = xus[e] - mean_u;
          Real cont = beta[KT - 1];
#pragma unroll
          for (int k = KT - 2; k >= 0; --k) cont = fma(cont, c, beta[k]);
          cont = cont > Real(0) ? cont : Real(0);
          wn[e] = ev > cont ? ev : ratio_u * wn[e];
        }
        V out;
        out.x = wn[0];
        out.y = wn[1];
        st_keep(w + p, out, keep);
      }
      if (A.do_acc) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const Real v = strike - xas[e];
          const bool use = v > Real(0) &&
                           (calib_all || (A.path_offset + 2ull * p + e) < A.num_calib);
          if (use) {
            double phi[KT];
            const Real cr = xas[e] - mean_a;
            Real pw = 1;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
              phi[k] = static_cast<double>(pw);
              pw *= cr;
            }
            const double y = static_cast<double>(ratio_a * wn[e]);
            int idx = 0;
#pragma unroll
            for (int i = 0; i < KT; ++i)
#pragma unroll
              for (int j = i; j < KT; ++j) {
                acc[idx] = fma(phi[i], phi[j], acc[idx]);
                ++idx;
              }
#pragma unroll
            for (int i = 0; i < KT; ++i)
              acc[KT * (KT + 1) / 2 + i] = fma(phi[i], y, acc[KT * (KT + 1) / 2 + i]);
          }
        }
      }
    }
  }
  if (A.do_acc) {
    __shared__ double s_red[kLsmBlock / 32][NA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLsmFastNS) {
      const int slot = threadIdx.x;
      int src = -1;
      if (slot < kLsmFastK * (kLsmFastK + 1) / 2) {
        int i = 0, rem = slot;
        while (rem >= kLsmFastK - i) {
          rem -= kLsmFastK - i;
          ++i;
        }
        const int j = i + rem;
        if (i < KT && j < KT) src = i * KT - i * (i - 1) / 2 + (j - i);
      } else {
        const int i = slot - kLsmFastK * (kLsmFastK + 1) / 2;
        if (i < KT) src = KT * (KT + 1) / 2 + i;
      }
      double v = 0.0;
      if (src >= 0)
        for (int wi = 0; wi < kLsmBlock / 32; ++wi) v += s_red[wi][src];
      A.partials[(static_cast<size_t>(blockIdx.x) * A.batch + b) * kLsmFastNS + slot] = v;
    }
    if (A.ticket != nullptr) {
      // last CTA done: reduce the per-CTA rows in a fixed order and solve -- the
      // separate solve launch (and its launch latency) disappears from the
      // date-to-date critical path
      __shared__ bool s_last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(A.ticket, 1u);
        s_last = t == gridDim.x * gridDim.y - 1;
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        PeerK pk;
        pk.peer_rank = A.peer_rank;
        pk.peer_world = A.peer_world;
        pk.peer_epoch = A.peer_epoch;
#pragma unroll
        for (int r = 0; r < kLsmMaxPeers; ++r) pk.peer_bufs[r] = A.peer_bufs[r];
        lsm_fused_tail<KT>(pk, A.partials, A.batch, gridDim.x, A.sums_out, A.rcond,
                           A.round_to_float, A.beta_out);
        if (threadIdx.x == 0) *A.ticket = 0u;
      }
    }
  }
}

// Generic path (K <= 128): the update is per thread, the outer products are
// formed tile by tile from shared memory.  grid = (blocks, B); NS = K*K + K.
template <typename Real>
__global__ void __launch_bounds__(kLsmBlock) lsm_step_generic_kernel(const LsmArgs<Real> A) {
  extern __shared__ double s_gen[];   // [kLsmTile][K] phi, [kLsmTile] y
  const int b = blockIdx.y;
  const int K = A.K;
  const Real* base = A.paths + b * A.stride_batch;
  Real* w = A.w + static_cast<size_t>(b) * A.num_paths;
  double* s_phi = s_gen;
  double* s_y = s_gen + kLsmTile * K;
  const int NS = K * K + K;
  double* out = A.partials + (static_cast<size_t>(blockIdx.x) * A.batch + b) * NS;
  for (int i = threadIdx.x; i < NS; i += blockDim.x) out[i] = 0.0;
  const Real ratio_u = A.do_update ? static_cast<Real>(A.ratio_update[b]) : Real(1);
  const Real ratio_a = A.do_acc ? static_cast<Real>(A.ratio_acc[b]) : Real(1);
  Real phi[kLsmMaxK];
  const uint64_t tiles = (A.num_paths + kLsmTile - 1) / kLsmTile;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < kLsmTile) {
      const uint64_t n = tile * kLsmTile + threadIdx.x;
      bool use = false;
      double y = 0.0;
      if (n < A.num_paths) {
        const Real* xn = base + static_cast<int64_t>(n) * A.stride_path;
        Real wn = w[n];
        if (A.do_update) {
          const Real* xp = xn + A.t_update * A.stride_time;
          const Real ev = lsm_payoff(A, xp, b);
          lsm_basis(A, xp, A.mean_update + b * A.mean_stride, phi);
          Real cont = 0;
          for (int k = 0; k < K; ++k) cont += phi[k] * static_cast<Real>(A.beta[b * K + k]);
          cont = cont > Real(0) ? cont : Real(0);
          wn = ev > cont ? ev : ratio_u * wn;
          w[n] = wn;
        }
        if (A.do_acc) {
          const Real* xp = xn + A.t_acc * A.stride_time;
          const Real ev = lsm_payoff(A, xp, b);
          use = ev > Real(0) && (A.path_offset + n) < A.num_calib;
          if (use) {
            lsm_basis(A, xp, A.mean_acc + b * A.mean_stride, phi);
            y = static_cast<double>(ratio_a * wn);
          }
        }
      }
      for (int k = 0; k < K; ++k) s_phi[threadIdx.x * K + k] = use ? static_cast<double>(phi[k]) : 0.0;
      s_y[threadIdx.x] = use ? y : 0.0;
    }
    __syncthreads();
    if (A.do_acc) {
      for (int e = threadIdx.x; e < NS; e += blockDim.x) {
        double v = 0.0;
        if (e < K * K) {
          const int i = e / K, j = e - i * K;
          for (int p = 0; p < kLsmTile; ++p) v += s_phi[p * K + i] * s_phi[p * K + j];
        } else {
          const int i = e - K * K;
          for (int p = 0; p < kLsmTile; ++p) v += s_phi[p * K + i] * s_y[p];
        }
        out[e] += v;
      }
    }
  }
}

template <typename Real>
__global__ void __launch_bounds__(kLsmBlock) lsm_init_kernel(const LsmArgs<Real> A, int t_index) {
  const int b = blockIdx.y;
  const Real* base = A.paths + b * A.stride_batch;
  Real* w = A.w + static_cast<size_t>(b) * A.num_paths;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t n = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < A.num_paths;
       n += stride) {
    const Real* xp = base + static_cast<int64_t>(n) * A.stride_path + t_index * A.stride_time;
    w[n] = A.ev_tab ? A.ev_tab[(static_cast<size_t>(A.slot_update) * A.batch + b) * A.num_paths + n]
                    : lsm_payoff(A, xp, b);
  }
}

// Column sums sum_n x[n, t, j] for the basis centring (lsm.py:110-111).
// grid = (blocks, num_columns, B); column c = (time slot, dim j).
template <typename Real>
__global__ void __launch_bounds__(kLsmBlock) lsm_colsum_kernel(const LsmArgs<Real> A,
                                                               const int* time_indices,
                                                               int num_times, double* partials) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int ti = c / A.dim, j = c - ti * A.dim;
  const Real* base = A.paths + b * A.stride_batch + time_indices[ti] * A.stride_time +
                     j * A.stride_dim;
  double s = 0.0;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t n = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < A.num_paths;
       n += stride)
    s += static_cast<double>(base[static_cast<int64_t>(n) * A.stride_path]);
  __shared__ double s_red[kLsmBlock / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int wi = 0; wi < kLsmBlock / 32; ++wi) v += s_red[wi];
    partials[(static_cast<size_t>(blockIdx.x) * A.batch + b) * (num_times * A.dim) + c] = v;
  }
}

// sum_n W[b][n] over the non-calibration paths -> partials [grid][B][2].
template <typename Real>
__global__ void __launch_bounds__(kLsmBlock) lsm_wsum_kernel(const LsmArgs<Real> A,
                                                             uint64_t skip_below,
                                                             double* partials) {
  const int b = blockIdx.y;
  const Real* w = A.w + static_cast<size_t>(b) * A.num_paths;
  double s = 0.0, cnt = 0.0;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t n = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < A.num_paths;
       n += stride) {
    if (A.path_offset + n >= skip_below) {
      s += static_cast<double>(A.ratio_path ? A.ratio_path[n] * w[n] : w[n]);
      cnt += 1.0;
    }
  }
  __shared__ double s_red[2][kLsmBlock / 32];
  s = warp_sum(s);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = s;
    s_red[1][threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double v = 0.0;
    for (int wi = 0; wi < kLsmBlock / 32; ++wi) v += s_red[threadIdx.x][wi];
    partials[(static_cast<size_t>(blockIdx.x) * A.batch + b) * 2 + threadIdx.x] = v;
  }
}

// sums[m] = sum_blocks partials[block][m]: one warp per output, lanes stride
// over the blocks, fixed order -> reproducible.
__global__ void lsm_reduce_kernel(const double* __restrict__ partials, int num_blocks, int M,
                                  double* __restrict__ sums) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int m = blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < M;
       m += gridDim.x * warps_per_block) {
    double v = 0.0;
    for (int bk = lane; bk < num_blocks; bk += 32) v += partials[static_cast<size_t>(bk) * M + m];
    v = warp_sum(v);
    if (lane == 0) sums[m] = v;
  }
}

// `partials` != nullptr: first reduce the per-CTA partial rows (fixed order)
// into `sums`, then solve -- one launch instead of two on a single GPU.
__global__ void lsm_solve_kernel(const double* __restrict__ partials, int num_blocks,
                                 double* __restrict__ sums, int B, int K, double rcond,
                                 int round_to_float, double* __restrict__ beta) {
  if (partials != nullptr) {
    const int M = B * kLsmFastNS;
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // one warp per sum, 8 warps (the 188 registers of the solver cap the block at 256 threads);
    // 8 independent loads in flight per lane, combined in a fixed order
    for (int m = threadIdx.x >> 5; m < M; m += nwarps) {
      double a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = 0.0;
      int bk = lane;
      for (; bk + 7 * 32 < num_blocks; bk += 8 * 32) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += partials[static_cast<size_t>(bk + u * 32) * M + m];
      }
      double tail = 0.0;
      for (; bk < num_blocks; bk += 32) tail += partials[static_cast<size_t>(bk) * M + m];
      double v = (((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))) + tail;
      v = warp_sum(v);
      if (lane == 0) sums[m] = v;
    }
    __syncthreads();
  }
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* sp = sums + static_cast<size_t>(b) * kLsmFastNS;
  double* out = beta + static_cast<size_t>(b) * K;
  switch (K) {
    case 1: lsm_solve_one<1>(sp, rcond, round_to_float, out); break;
    case 2: lsm_solve_one<2>(sp, rcond, round_to_float, out); break;
    case 3: lsm_solve_one<3>(sp, rcond, round_to_float, out); break;
    case 4: lsm_solve_one<4>(sp, rcond, round_to_float, out); break;
    case 5: lsm_solve_one<5>(sp, rcond, round_to_float, out); break;
    default: lsm_solve_one<6>(sp, rcond, round_to_float, out); break;
  }
}

}  // namespace tqf

using namespace tqf;

// Exercise-date slot of a time index (tabulated mode); -1 when unknown.
static int slot_of(const tqf_lsm* h, int time_index) {
  if (!h->exercise_times) return -1;
  for (size_t i = 0; i < h->exercise_times->size(); ++i)
    if ((*h->exercise_times)[i] == time_index) return static_cast<int>(i);
  return -1;
}

template <typename Real>
static void fill_args(const tqf_lsm* h, LsmArgs<Real>* A) {
  std::memset(A, 0, sizeof(*A));
  const tqf_lsm_desc& d = h->desc;
  A->paths = static_cast<const Real*>(d.paths_dev);
  A->stride_path = d.stride_path;
  A->stride_time = d.stride_time;
  A->stride_dim = d.stride_dim;
  A->stride_batch = d.stride_batch;
  A->w = static_cast<Real*>(h->w_dev);
  A->num_paths = d.num_paths;
  A->path_offset = d.path_offset;
  A->num_calib = d.num_calibration_samples == 0 ? ~0ull : d.num_calibration_samples;
  A->dim = d.dim;
  A->K = h->K;
  A->batch = d.batch;
  A->exponents = h->exponents_dev;
  A->strikes = h->strikes_dev;
  A->partials = h->partials_dev;
  A->NS = h->NS;
  A->ev_tab = static_cast<const Real*>(d.exercise_values_dev);
  A->ratio_path = static_cast<const Real*>(d.path_ratio_dev);
}

static int ensure_partials(tqf_lsm* h, size_t doubles) {
  if (doubles <= h->partials_doubles) return TQF_OK;
  if (h->external_partials) {
    set_error("caller-provided LSM partials workspace is too small");
    return TQF_ERR_INVALID_ARGUMENT;
  }
  {
    void* old = h->partials_dev;
    dev_release(&old, 1);
  }
  h->partials_dev = nullptr;
  h->partials_doubles = 0;
  void* fresh = nullptr;
  const int rc = dev_alloc(&fresh, doubles * sizeof(double));   // block cache (tqf_rng.cu)
  if (rc != TQF_OK) return rc;
  h->partials_dev = static_cast<double*>(fresh);
  h->partials_doubles = doubles;
  return TQF_OK;
}

template <typename Real>
static int lsm_step_impl(tqf_lsm* h, int do_update, int t_update, const double* mean_update,
                         const double* beta, const double* ratio_update, int do_acc, int t_acc,
                         const double* mean_acc, const double* ratio_acc, int64_t mean_stride,
                         double* sums_dev, cudaStream_t s) {
  const tqf_lsm_desc& d = h->desc;
  const int B = d.batch, K = h->K;
  LsmArgs<Real> A;
  fill_args(h, &A);
  A.do_update = do_update;
  A.t_update = t_update;
  A.mean_update = mean_update;
  A.beta = beta;
  A.ratio_update = ratio_update;
  A.do_acc = do_acc;
  A.t_acc = t_acc;
  A.mean_acc = do_acc ? mean_acc : mean_update;
  A.ratio_acc = do_acc ? ratio_acc : ratio_update;
  A.mean_stride = mean_stride;
  if (h->tabulated) {
    A.slot_update = do_update ? slot_of(h, t_update) : 0;
    A.slot_acc = do_acc ? slot_of(h, t_acc) : 0;
    TQF_REQUIRE(A.slot_update >= 0 && A.slot_acc >= 0,
                "time index is not one of exercise_time_indices");
  }
  int rc = ensure_partials(h, static_cast<size_t>(h->grid) * B * h->NS);
  if (rc != TQF_OK) return rc;
  A.partials = h->partials_dev;
  const dim3 grid(h->grid, B);
  const bool vec_ok = lsm_vec_ok(h);
  h->last_step_solved = false;
  if (vec_ok && do_acc && h->ticket_dev != nullptr) {
    A.ticket = h->ticket_dev;
    A.sums_out = h->fused_sums_dev;
    A.beta_out = h->fused_beta_dev;
    A.rcond = h->fused_rcond;
    A.round_to_float = d.dtype == TQF_F32 ? 1 : 0;
    h->last_step_solved = true;
    A.peer_rank = h->peer_rank;
    A.peer_world = h->peer_world;
    if (h->peer_world > 1) {
      A.peer_epoch = ++h->peer_epoch;
      for (int r = 0; r < h->peer_world; ++r) A.peer_bufs[r] = h->peer_bufs[r];
    }
  }
  if (vec_ok) {
    switch (K) {
      case 1: lsm_step_dim1_vec_kernel<Real, 1><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 2: lsm_step_dim1_vec_kernel<Real, 2><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 3: lsm_step_dim1_vec_kernel<Real, 3><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 4: lsm_step_dim1_vec_kernel<Real, 4><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 5: lsm_step_dim1_vec_kernel<Real, 5><<<grid, kLsmBlock, 0, s>>>(A); break;
      default: lsm_step_dim1_vec_kernel<Real, 6><<<grid, kLsmBlock, 0, s>>>(A); break;
    }
  } else if (h->fast && d.dim == 1 && !h->tabulated) {
    switch (K) {
      case 1: lsm_step_dim1_kernel<Real, 1><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 2: lsm_step_dim1_kernel<Real, 2><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 3: lsm_step_dim1_kernel<Real, 3><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 4: lsm_step_dim1_kernel<Real, 4><<<grid, kLsmBlock, 0, s>>>(A); break;
      case 5: lsm_step_dim1_kernel<Real, 5><<<grid, kLsmBlock, 0, s>>>(A); break;
      default: lsm_step_dim1_kernel<Real, 6><<<grid, kLsmBlock, 0, s>>>(A); break;
    }
  } else if (h->fast) {
    lsm_step_fast_kernel<Real><<<grid, kLsmBlock, 0, s>>>(A);
  } else {
    const size_t smem = (static_cast<size_t>(kLsmTile) * K + kLsmTile) * sizeof(double);
    lsm_step_generic_kernel<Real><<<grid, kLsmBlock, smem, s>>>(A);
  }
  TQF_CUDA_OK(cudaGetLastError());
  if (do_acc && sums_dev != nullptr) {
    const int M = B * h->NS;
    const int blocks = (M + 3) / 4 < 592 ? (M + 3) / 4 : 592;
    lsm_reduce_kernel<<<blocks, 128, 0, s>>>(h->partials_dev, h->grid, M, sums_dev);
    TQF_CUDA_OK(cudaGetLastError());
  }
  return TQF_OK;
}

extern "C" {

int tqf_lsm_workspace(const tqf_lsm_desc* desc, int num_times, uint64_t* partials_doubles) {
  TQF_REQUIRE(desc && partials_doubles && num_times >= 1, "bad arguments");
  const int K = desc->basis_size;
  const int NS = K <= kLsmFastK ? kLsmFastNS : K * K + K;
  int sms = kSMs;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess) {
    cudaGetLastError();
    sms = kSMs;
  }
  const uint64_t grid = static_cast<uint64_t>(sms) * 8;
  uint64_t need = grid * desc->batch * NS;
  const uint64_t cols = 64ull * desc->batch * num_times * desc->dim;
  if (cols > need) need = cols;
  *partials_doubles = need + 64;
  return TQF_OK;
}

int tqf_lsm_create(const tqf_lsm_desc* desc, tqf_lsm** out) {
  TQF_NVTX("tqf_lsm_create");
  TQF_REQUIRE(desc && out, "null argument");
  *out = nullptr;
  TQF_REQUIRE(desc->dtype == TQF_F32 || desc->dtype == TQF_F64, "bad dtype");
  TQF_REQUIRE(desc->dim >= 1 && desc->dim <= kLsmMaxDim, "dim must be in [1, 8]");
  TQF_REQUIRE(desc->batch >= 1, "batch must be >= 1");
  TQF_REQUIRE(desc->basis_size >= 1 && desc->basis_size <= kLsmMaxK,
              "basis size must be in [1, 128]");
  TQF_REQUIRE(desc->exponents && desc->strikes && desc->paths_dev, "null pointer in descriptor");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libtqf has no CPU fallback");
    return TQF_ERR_CUDA;
  }
  tqf_lsm* h = new (std::nothrow) tqf_lsm();
  TQF_REQUIRE(h, "out of memory");
  std::memset(h, 0, sizeof(*h));
  h->desc = *desc;
  h->K = desc->basis_size;
  // the register path assumes the 1-d exponent order 0..K-1 or takes any
  // exponents for dim > 1 through lsm_basis
  h->fast = h->K <= kLsmFastK;
  h->NS = h->fast ? kLsmFastNS : h->K * h->K + h->K;
  int sms = kSMs;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint64_t blocks = (desc->num_paths + kLsmBlock - 1) / kLsmBlock;
  // fast kernels: 3 CTAs of 256 threads resident per SM -> one full wave (and few partial rows)
  const uint64_t cap = static_cast<uint64_t>(sms) * (h->fast ? 3 : 2);
  h->grid = static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap);
  // the one-load-per-iteration kernels (initial payoff, value sum) want more CTAs in flight
  const uint64_t cap_aux = static_cast<uint64_t>(sms) * 8;
  h->grid_aux = static_cast<int>(blocks < cap_aux ? (blocks ? blocks : 1) : cap_aux);
  const size_t esize = desc->dtype == TQF_F64 ? 8 : 4;
  cudaError_t e = cudaSuccess;
  if (desc->w_dev) {               // caller-owned workspace (framework allocator)
    h->w_dev = desc->w_dev;
    h->external_w = true;
  } else {
    if (dev_alloc(&h->w_dev, esize * desc->batch * (desc->num_paths ? desc->num_paths : 1)) != TQF_OK)
      e = cudaErrorMemoryAllocation;
  }
  if (desc->partials_dev) {
    h->partials_dev = desc->partials_dev;
    h->partials_doubles = desc->partials_doubles;
    h->external_partials = true;
  }
  h->strike0 = desc->strikes[0];
  // (small per-object buffers: from the block cache, no cudaMalloc / cudaFree per pricing call)
  auto cached = [&e](auto** ptr, size_t bytes) {
    void* p = nullptr;
    if (e == cudaSuccess && dev_alloc(&p, bytes) != TQF_OK) e = cudaErrorMemoryAllocation;
    *ptr = static_cast<std::remove_reference_t<decltype(**ptr)>*>(p);
  };
  cached(&h->ctrl_dev, 4 * sizeof(unsigned long long));
  cached(&h->exponents_dev, sizeof(int) * h->K * desc->dim);
  cached(&h->strikes_dev, sizeof(double) * desc->batch);
  if (e == cudaSuccess)
    e = cudaMemcpy(h->exponents_dev, desc->exponents, sizeof(int) * h->K * desc->dim,
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(h->strikes_dev, desc->strikes, sizeof(double) * desc->batch,
                   cudaMemcpyHostToDevice);
  h->desc.exponents = nullptr;
  h->desc.strikes = nullptr;
  h->tabulated = desc->exercise_values_dev != nullptr || desc->path_ratio_dev != nullptr;
  if (h->tabulated && e == cudaSuccess) {
    if (!h->fast || !desc->exercise_time_indices || desc->num_exercise_times < 1) {
      tqf_lsm_destroy(h);
      set_error("tabulated exercise values / per-path ratios need K <= 6 and "
                "exercise_time_indices");
      return TQF_ERR_INVALID_ARGUMENT;
    }
    h->exercise_times = new std::vector<int>(
        desc->exercise_time_indices, desc->exercise_time_indices + desc->num_exercise_times);
  }
  h->desc.exercise_time_indices = nullptr;
  if (e != cudaSuccess) {
    tqf_lsm_destroy(h);
    return cuda_fail(e, "tqf_lsm_create");
  }
  *out = h;
  return TQF_OK;
}

int tqf_lsm_destroy(tqf_lsm* h) {
  if (!h) return TQF_OK;
  void* blocks[6] = {h->external_w ? nullptr : h->w_dev, h->exponents_dev, h->strikes_dev,
                     h->external_partials ? nullptr : static_cast<void*>(h->partials_dev),
                     h->times_dev, h->ctrl_dev};
  dev_release(blocks, 6);
  delete h->exercise_times;
  delete h;
  return TQF_OK;
}

int tqf_lsm_column_sums(tqf_lsm* h, const int32_t* time_indices, int num_times, double* sums_dev,
                        void* stream) {
  TQF_NVTX("tqf_lsm_column_sums");
  TQF_REQUIRE(h && time_indices && sums_dev && num_times >= 1, "bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const tqf_lsm_desc& d = h->desc;
  if (num_times > h->times_cap) {
    void* old = h->times_dev;
    dev_release(&old, 1);
    h->times_dev = nullptr;
    void* fresh = nullptr;
    const int rc = dev_alloc(&fresh, sizeof(int) * num_times);
    if (rc != TQF_OK) return rc;
    h->times_dev = static_cast<int*>(fresh);
    h->times_cap = num_times;
  }
  TQF_CUDA_OK(cudaMemcpyAsync(h->times_dev, time_indices, sizeof(int) * num_times,
                              cudaMemcpyHostToDevice, s));
  const int cols = num_times * d.dim;
  const int gx = h->grid < 64 ? h->grid : 64;
  int rc = ensure_partials(h, static_cast<size_t>(gx) * d.batch * cols);
  if (rc != TQF_OK) return rc;
  const dim3 grid(gx, cols, d.batch);
  if (d.dtype == TQF_F64) {
    LsmArgs<double> A;
    fill_args(h, &A);
    lsm_colsum_kernel<double><<<grid, kLsmBlock, 0, s>>>(A, h->times_dev, num_times, h->partials_dev);
  } else {
    LsmArgs<float> A;
    fill_args(h, &A);
    lsm_colsum_kernel<float><<<grid, kLsmBlock, 0, s>>>(A, h->times_dev, num_times, h->partials_dev);
  }
  TQF_CUDA_OK(cudaGetLastError());
  const int M = d.batch * cols;
  lsm_reduce_kernel<<<(M + 3) / 4 < 592 ? (M + 3) / 4 : 592, 128, 0, s>>>(h->partials_dev, gx, M, sums_dev);
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_lsm_init(tqf_lsm* h, int time_index, void* stream) {
  TQF_NVTX("tqf_lsm_init");
  TQF_REQUIRE(h, "null handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid(h->grid_aux, h->desc.batch);
  const int slot = h->tabulated ? slot_of(h, time_index) : 0;
  TQF_REQUIRE(slot >= 0, "time index is not one of exercise_time_indices");
  if (h->desc.dtype == TQF_F64) {
    LsmArgs<double> A;
    fill_args(h, &A);
    A.slot_update = slot;
    lsm_init_kernel<double><<<grid, kLsmBlock, 0, s>>>(A, time_index);
  } else {
    LsmArgs<float> A;
    fill_args(h, &A);
    A.slot_update = slot;
    lsm_init_kernel<float><<<grid, kLsmBlock, 0, s>>>(A, time_index);
  }
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_lsm_step(tqf_lsm* h, int do_update, int t_update, const double* mean_update_dev,
                 const double* beta_dev, const double* ratio_update_dev, int do_accumulate,
                 int t_acc, const double* mean_acc_dev, const double* ratio_acc_dev,
                 int64_t mean_stride, double* sums_dev, void* stream) {
  TQF_NVTX("tqf_lsm_step");
  TQF_REQUIRE(h, "null handle");
  TQF_REQUIRE(!do_update || (mean_update_dev && beta_dev && ratio_update_dev),
              "null update argument");
  TQF_REQUIRE(!do_accumulate || (mean_acc_dev && ratio_acc_dev), "null accumulate argument");
  TQF_REQUIRE(!do_accumulate || sums_dev || h->fast,
              "sums_dev may only be omitted for the packed (K <= 6) layout");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return h->desc.dtype == TQF_F64
             ? lsm_step_impl<double>(h, do_update, t_update, mean_update_dev, beta_dev,
                                     ratio_update_dev, do_accumulate, t_acc, mean_acc_dev,
                                     ratio_acc_dev, mean_stride, sums_dev, s)
             : lsm_step_impl<float>(h, do_update, t_update, mean_update_dev, beta_dev,
                                    ratio_update_dev, do_accumulate, t_acc, mean_acc_dev,
                                    ratio_acc_dev, mean_stride, sums_dev, s);
}

int tqf_lsm_solve(tqf_lsm* h, double* sums_dev, int reduce_partials, double rcond,
                  double* beta_dev, void* stream) {
  TQF_NVTX("tqf_lsm_solve");
  TQF_REQUIRE(h && sums_dev && beta_dev, "null argument");
  if (!h->fast) {
    set_error("device solve is implemented for basis sizes <= 6; solve on the host");
    return TQF_ERR_UNSUPPORTED;
  }
  if (h->last_step_solved && beta_dev == h->fused_beta_dev) {
    h->last_step_solved = false;   // the step kernel's last CTA already wrote beta
    return TQF_OK;
  }
  const int B = h->desc.batch;
  TQF_REQUIRE(!reduce_partials || B <= 128, "fused reduction supports up to 128 payoffs");
  if (reduce_partials) {
    lsm_solve_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        h->partials_dev, h->grid, sums_dev, B, h->K, rcond, h->desc.dtype == TQF_F32 ? 1 : 0,
        beta_dev);
  } else {
    lsm_solve_kernel<<<(B + 31) / 32, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        nullptr, 0, sums_dev, B, h->K, rcond, h->desc.dtype == TQF_F32 ? 1 : 0, beta_dev);
  }
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_lsm_set_fused_solve(tqf_lsm* h, double rcond, double* sums_dev, double* beta_dev,
                            uint32_t* ticket_dev) {
  TQF_REQUIRE(h && sums_dev && beta_dev && ticket_dev, "null argument");
  if (!h->fast) {
    set_error("the fused solve is implemented for basis sizes <= 6");
    return TQF_ERR_UNSUPPORTED;
  }
  h->ticket_dev = ticket_dev;
  h->fused_sums_dev = sums_dev;
  h->fused_beta_dev = beta_dev;
  h->fused_rcond = rcond;
  return TQF_OK;
}

int tqf_lsm_run_fused(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                      const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                      double* beta_dev, void* stream) {
  TQF_NVTX("tqf_lsm_run_fused");
  TQF_REQUIRE(h && exercise_times && means_dev && ratio_dev && beta_dev && num_times >= 1,
              "bad arguments");
  TQF_REQUIRE(h->ticket_dev != nullptr && beta_dev == h->fused_beta_dev && lsm_vec_ok(h),
              "tqf_lsm_run_fused needs tqf_lsm_set_fused_solve on an eligible handle");
  const int dim = h->desc.dim, B = h->desc.batch;
  // exercise index e uses the means of time slot e - 1 and the ratio row e
  auto mean_of = [&](int e) { return means_dev + static_cast<int64_t>(e - 1) * dim; };
  auto ratio_of = [&](int e) { return ratio_dev + static_cast<int64_t>(e) * B; };
  int e = num_times - 1;
  int rc = TQF_OK;
  if (e > 0)
    rc = tqf_lsm_step(h, 0, 0, nullptr, nullptr, nullptr, 1, exercise_times[e - 1], mean_of(e),
                      ratio_of(e), mean_stride, nullptr, stream);
  while (rc == TQF_OK && e > 0) {
    const int do_acc = e - 1 > 0 ? 1 : 0;
    rc = tqf_lsm_step(h, 1, exercise_times[e - 1], mean_of(e), beta_dev, ratio_of(e), do_acc,
                      do_acc ? exercise_times[e - 2] : 0, do_acc ? mean_of(e - 1) : nullptr,
                      do_acc ? ratio_of(e - 1) : nullptr, mean_stride, nullptr, stream);
    --e;
  }
  h->last_step_solved = false;
  return rc;
}

int tqf_lsm_persistent_eligible(const tqf_lsm* h, int* eligible) {
  TQF_REQUIRE(h && eligible, "null argument");
  *eligible = lsm_persistent_ok(h) ? 1 : 0;
  return TQF_OK;
}

int tqf_lsm_run_persistent(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                           const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                           double rcond, uint64_t skip_below, double* value_sums_dev,
                           double* beta_dev, double* history_dev, void* stream) {
  TQF_NVTX("tqf_lsm_run_persistent");
  TQF_REQUIRE(h && exercise_times && means_dev && ratio_dev && value_sums_dev && beta_dev &&
                  num_times >= 1,
              "bad arguments");
  for (int i = 0; i < num_times; ++i)
    TQF_REQUIRE(exercise_times[i] >= 0, "negative exercise time index");
  return lsm_run_persistent(h, exercise_times, num_times, means_dev, mean_stride, ratio_dev, rcond,
                            skip_below, value_sums_dev, beta_dev, history_dev,
                            static_cast<cudaStream_t>(stream));
}

int tqf_lsm_status(const tqf_lsm* h, uint64_t* status) {
  TQF_REQUIRE(h && status, "null argument");
  unsigned long long v = 0;
  if (h->ctrl_dev) TQF_CUDA_OK(cudaMemcpy(&v, h->ctrl_dev + 2, sizeof(v), cudaMemcpyDeviceToHost));
  *status = v;
  return TQF_OK;
}

int tqf_lsm_fused_eligible(const tqf_lsm* h, int* eligible) {
  TQF_REQUIRE(h && eligible, "null argument");
  *eligible = lsm_vec_ok(h) ? 1 : 0;
  return TQF_OK;
}

int tqf_lsm_peer_bytes(uint64_t* bytes) {
  TQF_REQUIRE(bytes, "null argument");
  *bytes = kLsmPeerBytes;
  return TQF_OK;
}

int tqf_lsm_set_peer_exchange(tqf_lsm* h, int rank, int world, void* const* bufs,
                              uint64_t epoch_base) {
  TQF_REQUIRE(h && bufs, "null argument");
  TQF_REQUIRE(world >= 1 && world <= kLsmMaxPeers && rank >= 0 && rank < world,
              "peer exchange supports up to 8 ranks");
  TQF_REQUIRE(h->desc.batch <= kLsmPeerMaxBatch, "peer exchange supports up to 16 payoffs");
  TQF_REQUIRE(lsm_vec_ok(h), "the fused pass does not apply to this problem (tqf_lsm_fused_eligible)");
  h->peer_rank = rank;
  h->peer_world = world;
  h->peer_epoch = epoch_base;
  for (int r = 0; r < world; ++r) {
    TQF_REQUIRE(bufs[r] != nullptr, "null peer buffer");
    h->peer_bufs[r] = static_cast<unsigned char*>(bufs[r]);
  }
  return TQF_OK;
}

int tqf_lsm_peer_epoch(const tqf_lsm* h, uint64_t* epoch) {
  TQF_REQUIRE(h && epoch, "null argument");
  *epoch = h->peer_epoch;
  return TQF_OK;
}

/* Peer-visible device memory of one box (CUDA IPC). */
int tqf_peer_alloc(uint64_t bytes, void** dev_ptr, uint8_t ipc_handle[64]) {
  TQF_REQUIRE(dev_ptr && ipc_handle && bytes > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  void* p = nullptr;
  TQF_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t hnd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hnd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "tqf_peer_alloc");
  }
  std::memcpy(ipc_handle, &hnd, 64);
  *dev_ptr = p;
  return TQF_OK;
}

int tqf_peer_open(const uint8_t ipc_handle[64], void** dev_ptr) {
  TQF_REQUIRE(dev_ptr && ipc_handle, "null argument");
  cudaIpcMemHandle_t hnd;
  std::memcpy(&hnd, ipc_handle, 64);
  TQF_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
  return TQF_OK;
}

int tqf_peer_close(void* dev_ptr) {
  if (dev_ptr) TQF_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return TQF_OK;
}

int tqf_peer_status(const void* own_buf, uint64_t* timeouts, uint64_t* last_epoch) {
  TQF_REQUIRE(own_buf && timeouts && last_epoch, "null argument");
  unsigned long long w[2] = {0, 0};
  TQF_CUDA_OK(cudaMemcpy(w, static_cast<const unsigned char*>(own_buf) + kLsmMaxPeers * 8, sizeof(w),
                         cudaMemcpyDeviceToHost));
  *timeouts = w[0];
  *last_epoch = w[1];
  return TQF_OK;
}

int tqf_peer_free(void* dev_ptr) {
  if (dev_ptr) TQF_CUDA_OK(cudaFree(dev_ptr));
  return TQF_OK;
}

int tqf_lsm_sums_layout(const tqf_lsm* h, int* num_sums, int* is_packed_symmetric) {
  TQF_REQUIRE(h && num_sums && is_packed_symmetric, "null argument");
  *num_sums = h->NS;
  *is_packed_symmetric = h->fast ? 1 : 0;
  return TQF_OK;
}

int tqf_lsm_value_sum(tqf_lsm* h, uint64_t skip_below, double* sums_dev, void* stream) {
  TQF_NVTX("tqf_lsm_value_sum");
  TQF_REQUIRE(h && sums_dev, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const tqf_lsm_desc& d = h->desc;
  int rc = ensure_partials(h, static_cast<size_t>(h->grid_aux) * d.batch * 2);
  if (rc != TQF_OK) return rc;
  const dim3 grid(h->grid_aux, d.batch);
  if (d.dtype == TQF_F64) {
    LsmArgs<double> A;
    fill_args(h, &A);
    lsm_wsum_kernel<double><<<grid, kLsmBlock, 0, s>>>(A, skip_below, h->partials_dev);
  } else {
    LsmArgs<float> A;
    fill_args(h, &A);
    lsm_wsum_kernel<float><<<grid, kLsmBlock, 0, s>>>(A, skip_below, h->partials_dev);
  }
  TQF_CUDA_OK(cudaGetLastError());
  const int M = d.batch * 2;
  lsm_reduce_kernel<<<1, 128, 0, s>>>(h->partials_dev, h->grid_aux, M, sums_dev);
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

}  // extern "C"
