// Correlated multi-asset geometric Brownian motion, Euler scheme (config C4:
// 64 assets, float32, Sobol).  One thread carries one path with all `dim`
// asset prices in registers; per step it draws `dim` normals in-kernel,
// applies the Cholesky factor and updates the state.
//
// Replaces, for the closures of
// models/geometric_brownian_motion/multivariate_geometric_brownian_motion.py:130-151,
// the reference's per-step [N, dim, dim] volatility tensor (16 KB per path and
// step for dim = 64, with tf.linalg.cholesky re-run every step, line 147) and
// the tf.linalg.matvec over it (models/euler_sampling.py:529):
//   x_i' = (x_i + dt mu_i x_i) + (sigma_i x_i) sqrt_dt sum_{j<=i} L_ij z_j.
//
// The Cholesky factor lives in the kernel parameter space (constant bank): the
// fully unrolled lower-triangular mat-vec reads every L_ij as an FFMA / DFMA
// constant operand -- no shared-memory traffic, no per-path matrix.  CUDA-core
// FP32; a tensor-core formulation would need the [paths x dim] normal tile in
// shared memory and only pays for the 48 % of the step that is the mat-vec.
#include <cstring>
#include <mutex>
#include <vector>

#include "tqf_paths_kernel.cuh"

namespace tqf {

template <typename Real, int DMAX>
struct MvParams {
  int dim, num_steps, num_steps_total, rngk;
  const Real* coef;  // device [S][2]: dt, sqrt_dt
  PhiloxKey key;
  PhiloxCtr ctr;
  const uint32_t* sobol_v;
  const double* logtab;  // device-global log table (read through L1)
  uint64_t first_index, path_offset, path_count, num_chunks, chunk_base;
  int mode, num_payoffs;
  PayoffK pay[TQF_MAX_PAYOFFS];
  double* partials;
  const int* record_slot;  // device [S+1]: eval flags (price) / slots (paths)
  Real* out;
  int64_t stride_path, stride_time, stride_dim;
  int store_exp;
  int exact_log;  // additive log-space step (exact sampler) instead of the Euler step
  Real x0[DMAX], mu[DMAX], sigma[DMAX];
  Real L[DMAX * (DMAX + 1) / 2];  // packed rows of the lower-triangular factor
};

template <typename Real, int DMAX>
__global__ void __launch_bounds__(kBlock)
mvgbm_kernel(const __grid_constant__ MvParams<Real, DMAX> P) {
  __shared__ uint32_t s_high[kSobolTileDims];
  __shared__ uint4 s_low[kSobolTileDims * 2];
  __shared__ double s_acc[kWarps * TQF_MAX_PAYOFFS * 3];
  // normals of one step, [dim][kBlock]: the draw loop is ROLLED (small code) and
  // hands its results to the unrolled mat-vec through shared memory -- the fully
  // unrolled version overflowed the instruction cache (ncu: 2.0 no_instruction
  // stalls per issue).
  extern __shared__ __align__(16) unsigned char s_dyn[];
  Real* s_z = reinterpret_cast<Real*>(s_dyn);
  const int tid = threadIdx.x;
  const fm::ConstTab tab(P.logtab);
  for (int i = tid; i < kWarps * TQF_MAX_PAYOFFS * 3; i += kBlock) s_acc[i] = 0.0;
  __syncthreads();

  uint32_t lowmask[kLowBits];
#pragma unroll
  for (int b = 0; b < kLowBits; ++b) {
    lowmask[b] = 0u - ((static_cast<uint32_t>(tid) >> b) & 1u);
    asm volatile("" : "+r"(lowmask[b]));
  }
  const int dim = P.dim;
  const int tile_steps = dim >= kSobolTileDims ? 1 : kSobolTileDims / dim;
  const uint64_t stream_stride = static_cast<uint64_t>(P.num_steps_total) * dim;

  for (uint64_t chunk = blockIdx.x; chunk < P.num_chunks; chunk += gridDim.x) {
    const uint64_t index = P.chunk_base + chunk * kBlock + tid;
    const bool valid = index >= P.first_index && index < P.first_index + P.path_count;
    const uint64_t local = index - P.first_index;
    Real x[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) x[i] = P.x0[i];
    PhiloxStreamV<Real, 1> stream;
    if (P.rngk == RNGK_PHILOX) {
      const uint64_t fe[1] = {valid ? (P.path_offset + local) * stream_stride : 0};
      stream.init(P.key, P.ctr, tab, fe);
    }

    auto eval_payoffs = [&](int step_index) {
      const int warp = tid >> 5, lane = tid & 31;
      Real m = 0;
#pragma unroll
      for (int i = 0; i < DMAX; ++i)
        if (i < dim) m += x[i];
      m = m / static_cast<Real>(dim);
      for (int q = 0; q < P.num_payoffs; ++q) {
        const PayoffK& d = P.pay[q];
        if (d.step != step_index) continue;
        Real xf = m;
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
          if (i == d.component) xf = x[i];
        double sum = 0.0, sq = 0.0, bad = 0.0;
        if (valid) {
          const double v = eval_payoff(d, static_cast<double>(xf), 0.0, 0.0);
          if (isfinite(v)) {
            sum = v;
            sq = v * v;
          } else {
            bad = 1.0;
          }
        }
        sum = warp_sum(sum);
        sq = warp_sum(sq);
        bad = warp_sum(bad);
        if (lane == 0) {
          double* acc = s_acc + (warp * TQF_MAX_PAYOFFS + q) * 3;
          acc[0] += sum;
          acc[1] += sq;
          acc[2] += bad;
        }
      }
    };
    auto store_state = [&](int slot) {
      if (valid) {
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
          if (i < dim)
            P.out[static_cast<int64_t>(local) * P.stride_path + slot * P.stride_time +
                  i * P.stride_dim] = P.store_exp ? static_cast<Real>(exp(x[i])) : x[i];
      }
    };
    if (P.record_slot[0] >= 0) {
      if (P.mode == MODE_PRICE) eval_payoffs(0); else store_state(P.record_slot[0]);
    }

    for (int s0 = 0; s0 < P.num_steps; s0 += tile_steps) {
      const int s1 = min(P.num_steps, s0 + tile_steps);
      if (P.rngk == RNGK_SOBOL) {
        __syncthreads();
        const uint32_t high_bits = static_cast<uint32_t>((P.chunk_base + chunk * kBlock) >> kLowBits);
        for (int dd = tid; dd < (s1 - s0) * dim; dd += kBlock) {
          const uint32_t* v = P.sobol_v + (static_cast<size_t>(s0) * dim + dd) * 32;
          s_low[2 * dd] = *reinterpret_cast<const uint4*>(v);
          s_low[2 * dd + 1] = *reinterpret_cast<const uint4*>(v + 4);
          uint32_t hb = high_bits, h = 0;
          while (hb) {
            const int b = __ffs(hb) - 1;
            h ^= v[kLowBits + b];
            hb &= hb - 1;
          }
          s_high[dd] = h;
        }
        __syncthreads();
      }
      for (int s = s0; s < s1; ++s) {
        if (P.rngk == RNGK_SOBOL) {
          const uint4* lp = s_low + 2 * (s - s0) * dim;
          const uint32_t* hp = s_high + (s - s0) * dim;
#pragma unroll 8
          for (int j = 0; j < dim; ++j) {
            const uint4 l0 = lp[2 * j];
            const uint4 l1 = lp[2 * j + 1];
            uint32_t xb = hp[j];
            xb ^= l0.x & lowmask[0];
            xb ^= l0.y & lowmask[1];
            xb ^= l0.z & lowmask[2];
            xb ^= l0.w & lowmask[3];
            xb ^= l1.x & lowmask[4];
            xb ^= l1.y & lowmask[5];
            xb ^= l1.z & lowmask[6];
            const uint32_t xin[1] = {xb};
            Real zo[1];
            sobol_normals<1>(tab, xin, zo);
            s_z[j * kBlock + tid] = zo[0];
          }
        } else {
          for (int j = 0; j < dim; ++j) {
            Real zo[1];
            stream.next(P.key, P.ctr, tab, zo);
            s_z[j * kBlock + tid] = zo[0];
          }
        }
        // own column only: no barrier needed between the writes above and the reads
        Real z[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; ++j) z[j] = j < dim ? s_z[j * kBlock + tid] : Real(0);
        // correlate in place, highest row first: z_i <- sum_{j<=i} L_ij z_j
        const Real dt = P.coef[2 * s], sq = P.coef[2 * s + 1];
        // Rows are processed from the bottom in blocks of RB, each block with RB
        // independent accumulators (the rows only read z_j, j <= i, which are
        // still the raw normals: results overwrite z from the top down).
        constexpr int RB = DMAX >= 8 ? 8 : DMAX;
#pragma unroll
        for (int ib = DMAX - 1; ib >= 0; ib -= RB) {
          Real acc[RB];
#pragma unroll
          for (int r = 0; r < RB; ++r) acc[r] = 0;
#pragma unroll
          for (int j = 0; j <= ib; ++j) {
#pragma unroll
            for (int r = 0; r < RB; ++r) {
              const int i = ib - r;
              if (i >= 0 && j <= i) acc[r] = fma(P.L[i * (i + 1) / 2 + j], z[j], acc[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int i = ib - r;
            if (i >= 0) {
              z[i] = acc[r];
              if (P.exact_log) {
                // exact log-normal increment (multivariate_geometric_brownian_motion.py:262-266);
                // mu holds means - vols^2 / 2
                x[i] = x[i] + (P.mu[i] * dt + (sq * P.sigma[i]) * z[i]);
              } else {
                const Real dt_inc = dt * (P.mu[i] * x[i]);
                const Real dw_inc = (P.sigma[i] * x[i]) * (z[i] * sq);
                x[i] = (x[i] + dt_inc) + dw_inc;
              }
            }
          }
        }
        const int flag = P.record_slot[s + 1];
        if (flag >= 0) {
          if (P.mode == MODE_PRICE) eval_payoffs(s + 1); else store_state(flag);
        }
      }
    }
  }
  if (P.mode == MODE_PRICE) {
    __syncthreads();
    for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += kBlock) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += s_acc[w * TQF_MAX_PAYOFFS * 3 + i];
      const int q = i / 3, k = i - q * 3;
      P.partials[(static_cast<size_t>(blockIdx.x) * TQF_MAX_PAYOFFS + q) * 4 + k] = v;
    }
  }
}

template <typename Real, int DMAX>
static int launch_mv(const MvLaunch& a, cudaStream_t stream, int* grid_out) {
  static MvParams<Real, DMAX> P;   // large (up to 19 KB): keep off the stack
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  std::memset(&P, 0, sizeof(P));
  P.dim = a.dim;
  P.num_steps = a.num_steps;
  P.num_steps_total = a.num_steps_total;
  P.rngk = a.rngk;
  P.coef = static_cast<const Real*>(a.coef_dev);
  P.key = a.key;
  P.ctr = a.ctr;
  P.sobol_v = a.sobol_v;
  P.logtab = a.logtab;
  P.first_index = a.first_index;
  P.path_offset = a.path_offset;
  P.path_count = a.path_count;
  P.chunk_base = a.first_index & ~static_cast<uint64_t>(kBlock - 1);
  P.num_chunks = (a.first_index + a.path_count - P.chunk_base + kBlock - 1) / kBlock;
  P.mode = a.mode;
  P.num_payoffs = a.num_payoffs;
  for (int q = 0; q < a.num_payoffs; ++q) P.pay[q] = a.pay[q];
  P.partials = a.partials;
  P.record_slot = a.record_dev;
  P.out = static_cast<Real*>(a.out);
  P.stride_path = a.stride_path;
  P.stride_time = a.stride_time;
  P.stride_dim = a.stride_dim;
  P.store_exp = a.store_exp;
  P.exact_log = a.exact_log;
  for (int i = 0; i < DMAX; ++i) {
    const bool in = i < a.dim;
    P.x0[i] = in ? static_cast<Real>(a.x0[i]) : Real(0);
    P.mu[i] = in ? static_cast<Real>(a.mu[i]) : Real(0);
    P.sigma[i] = in ? static_cast<Real>(a.sigma[i]) : Real(0);
    for (int j = 0; j <= i; ++j)
      P.L[i * (i + 1) / 2 + j] =
          (in && j < a.dim) ? static_cast<Real>(a.chol[static_cast<size_t>(i) * a.dim + j]) : Real(0);
  }
  int grid = static_cast<int>(P.num_chunks < static_cast<uint64_t>(a.max_grid)
                                  ? P.num_chunks
                                  : static_cast<uint64_t>(a.max_grid));
  if (grid < 1) grid = 1;
  *grid_out = grid;
  const size_t smem = static_cast<size_t>(DMAX) * kBlock * sizeof(Real);
  if (smem > 48 * 1024)
    TQF_CUDA_OK(cudaFuncSetAttribute(mvgbm_kernel<Real, DMAX>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  mvgbm_kernel<Real, DMAX><<<grid, kBlock, smem, stream>>>(P);
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int launch_mvgbm(const MvLaunch& a, cudaStream_t stream, int* grid_out) {
  const bool f64 = a.dtype == TQF_F64;
  if (a.dim <= 8) return f64 ? launch_mv<double, 8>(a, stream, grid_out) : launch_mv<float, 8>(a, stream, grid_out);
  if (a.dim <= 64) return f64 ? launch_mv<double, 64>(a, stream, grid_out) : launch_mv<float, 64>(a, stream, grid_out);
  set_error("MVGBM supports at most 64 assets");
  return TQF_ERR_UNSUPPORTED;
}

}  // namespace tqf
