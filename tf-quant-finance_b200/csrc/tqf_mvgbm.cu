// Correlated multi-asset geometric Brownian motion, Euler scheme (config C4:
// 64 assets, float32, Sobol).  Three kernels:
//   mvgbm_kernel        dim <= 8: one thread carries one path, the factor in the
//                       kernel parameter space;
//   mvgbm_mma_kernel    8 < dim <= 64, float32, Sobol: a warp carries 16 paths,
//                       the triangular mat-vec runs on the tensor cores
//                       (mma.sync TF32 with split operands), normals are drawn
//                       directly in the B-fragment layout;
//   mvgbm_split_kernel  8 < dim <= 64 otherwise (float64, Philox): four threads
//                       per path, factor rows streamed from shared memory.
// Per step they draw `dim` normals in-kernel, apply the Cholesky factor and
// update the state.
//
// Replaces, for the closures of
// models/geometric_brownian_motion/multivariate_geometric_brownian_motion.py:130-151,
// the reference's per-step [N, dim, dim] volatility tensor (16 KB per path and
// step for dim = 64, with tf.linalg.cholesky re-run every step, line 147) and
// the tf.linalg.matvec over it (models/euler_sampling.py:529):
//   x_i' = (x_i + dt mu_i x_i) + (sigma_i x_i) sqrt_dt sum_{j<=i} L_ij z_j.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
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
  const float* ndtab;    // device-global cubic table of the float32 inverse CDF (mma kernel)
  const Real* lsplit;    // device: factor / mu / sigma in the split kernel's order (dim > 8)
  uint64_t first_index, path_offset, path_count, num_chunks, chunk_base;
  int mode, num_payoffs;
  PayoffK pay[TQF_MAX_PAYOFFS];
  double* partials;
  const int* record_slot;  // device [S+1]: eval flags (price) / slots (paths)
  Real* out;
  int64_t stride_path, stride_time, stride_dim;
  int store_exp;
  int exact_log;  // additive log-space step (exact sampler) instead of the Euler step
  int sobol_clamp;  // float32 Sobol: u == 1.0 -> largest float below 1 (tqf_plan_set_sobol_clamp)
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
  // the Philox stream's Box-Muller logarithm reads the MID-free table stored behind it
  const fm::ConstTab tab0(P.logtab + 2 * TQF_LOGTAB_COUNT);
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
      stream.init(P.key, P.ctr, tab0, fe);
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
            sobol_normals<1>(tab, xin, zo, P.sobol_clamp);
            s_z[j * kBlock + tid] = zo[0];
          }
        } else {
          for (int j = 0; j < dim; ++j) {
            Real zo[1];
            stream.next(P.key, P.ctr, tab0, zo);
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


// ---------------------------------------------------------------------------
// dim > 8: one path is carried by kMvParts = 4 threads (one per warp of the
// CTA), each owning the 16 interleaved rows i = 4 r + part of the state and of
// the Cholesky factor.  The one-thread-per-path kernel above needs ~190
// registers for 64 assets (2 warps per scheduler, issue slots half empty); here
// a thread holds 16 prices + 16 accumulators, 8-10 CTAs fit on an SM and the
// FFMA stream of one warp covers the latencies of the others.
//   * lanes = 32 consecutive paths / Sobol indices (5 low index bits per lane);
//   * per step every thread draws 16 of the 64 normals (contiguous dimensions
//     16 part .. 16 part + 15) into shared memory, grouped by 4 so that the
//     mat-vec reads them back with 16-byte loads;
//   * the normals are double-buffered over steps: ONE barrier per step;
//   * row i needs z_0..z_i: the interleaved row assignment balances the
//     triangular work (496..544 FMAs per thread).
constexpr int kMvParts = 4, kMvRows = 16, kMvLanes = 32, kMvDim = 64;
constexpr int kMvTileDims = 256;   // Sobol dimensions staged at once (4 steps of 64)
constexpr int kMvLowBits = 5;

// Per part: 544 factor entries in the order mv_rows consumes them, then
// mu[16], sigma[16] of its rows.  The sm_100a ALU takes no constant-bank
// operand: coefficients fetched one LDCU each from the 8 KB parameter block miss
// the small constant cache (measured 80 cycles per LDCU) -- broadcast LDS.128
// from shared memory delivers four at a time.
constexpr int kMvSplitL = 544;                       // sum_r (4 r + 4)
constexpr int kMvSplitStride = kMvSplitL + 2 * kMvRows;

// 16-byte / 4-byte shared-memory accesses by shared-window address (no generic
// address conversion in the loops).
template <typename Real>
__device__ __forceinline__ void lds16(uint32_t addr, Real* dst) {
  uint32_t a, b, c, d;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
  if (sizeof(Real) == 4) {
    dst[0] = static_cast<Real>(__uint_as_float(a));
    dst[1] = static_cast<Real>(__uint_as_float(b));
    dst[2] = static_cast<Real>(__uint_as_float(c));
    dst[3] = static_cast<Real>(__uint_as_float(d));
  } else {
    dst[0] = static_cast<Real>(__hiloint2double(static_cast<int>(b), static_cast<int>(a)));
    dst[1] = static_cast<Real>(__hiloint2double(static_cast<int>(d), static_cast<int>(c)));
  }
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 q;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
  return q;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_real(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_real(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory");
}

// Rows i = 4 r + part of  x' = step(x, L z)  for the NP paths of a thread.
// Every part runs the SAME code (one copy in the instruction cache for all
// warps): row r is given the 4 r + 4 coefficients j = 0 .. 4 r + 3, those with
// j > i stored as zeros (4.6 % more FMAs than the exact triangle).  A
// coefficient costs one quarter of a broadcast LDS.128 and feeds NP FMAs: the
// shared-memory return path (512 B per warp-wide 16-byte load) is what bounds
// this kernel, so carrying two paths per thread halves its cost per path.
// `lp_addr`: this part's table, `z_addr`: this lane's normals of path 0 (path
// q is kMvLanes * 4 values further), both shared-window addresses.
template <typename Real, int NP>
__device__ __forceinline__ void mv_rows(uint32_t lp_addr, int exact_log, uint32_t z_addr, Real dt,
                                        Real sq, Real (&x)[NP][kMvRows]) {
  constexpr int PER16 = 16 / sizeof(Real);   // values per 16-byte load
  Real acc[NP][kMvRows];
#pragma unroll
  for (int q = 0; q < NP; ++q)
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) acc[q][r] = 0;
  Real cq[PER16];
  int k = 0;   // compile-time after unrolling
#pragma unroll
  for (int jg = 0; jg < kMvDim / 4; ++jg) {
    Real zq[NP][4];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
      for (int v = 0; v < 4 / PER16; ++v)
        lds16<Real>(z_addr + ((jg * NP + q) * kMvLanes * 4 + v * PER16) * sizeof(Real),
                    zq[q] + v * PER16);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int r = jg; r < kMvRows; ++r) {       // j = 4 jg + u <= 4 r + 3
        if ((k % PER16) == 0) lds16<Real>(lp_addr + k * sizeof(Real), cq);
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q][r] = fma(cq[k % PER16], zq[q][u], acc[q][r]);
        ++k;
      }
    }
  }
  Real mu[kMvRows], sg[kMvRows];
#pragma unroll
  for (int r = 0; r < kMvRows; r += PER16) {
    lds16<Real>(lp_addr + (kMvSplitL + r) * sizeof(Real), mu + r);
    lds16<Real>(lp_addr + (kMvSplitL + kMvRows + r) * sizeof(Real), sg + r);
  }
#pragma unroll
  for (int q = 0; q < NP; ++q)
#pragma unroll
    for (int r = 0; r < kMvRows; ++r) {
      if (exact_log) {
        // exact log-normal increment (multivariate_geometric_brownian_motion.py:262-266);
        // mu holds means - vols^2 / 2
        x[q][r] = x[q][r] + (mu[r] * dt + (sq * sg[r]) * acc[q][r]);
      } else {
        const Real dt_inc = dt * (mu[r] * x[q][r]);
        const Real dw_inc = (sg[r] * x[q][r]) * (acc[q][r] * sq);
        x[q][r] = (x[q][r] + dt_inc) + dw_inc;
      }
    }
}

// Host: factor (row-major [dim][dim], lower triangle), mu, sigma -> the split
// kernel's consumption order, [kMvParts][kMvSplitStride] values of `Real`.
template <typename Real>
static void build_split(const double* chol, const double* mu, const double* sigma, int dim,
                        std::vector<Real>* out) {
  out->assign(static_cast<size_t>(kMvParts) * kMvSplitStride, Real(0));
  for (int part = 0; part < kMvParts; ++part) {
    Real* lp = out->data() + static_cast<size_t>(part) * kMvSplitStride;
    int k = 0;
    for (int j = 0; j < kMvDim; ++j)
      for (int r = j / 4; r < kMvRows; ++r) {       // the order mv_rows reads them
        const int i = 4 * r + part;
        lp[k++] = (j <= i && i < dim && j < dim)
                      ? static_cast<Real>(chol[static_cast<size_t>(i) * dim + j]) : Real(0);
      }
    for (int r = 0; r < kMvRows; ++r) {
      const int i = 4 * r + part;
      lp[kMvSplitL + r] = i < dim ? static_cast<Real>(mu[i]) : Real(0);
      lp[kMvSplitL + kMvRows + r] = i < dim ? static_cast<Real>(sigma[i]) : Real(0);
    }
  }
}

static void build_mma(const double* chol, const double* mu, const double* sigma, int dim,
                      std::vector<float>* out);
static void build_tc5(const double* chol, const double* mu, const double* sigma, int dim,
                      std::vector<float>* out);

int mvgbm_upload_split(const double* chol, const double* mu, const double* sigma, int dim,
                       int dtype, void** out_dev) {
  void* dev = nullptr;
  cudaError_t e;
  if (dtype == TQF_F64) {
    std::vector<double> host;
    build_split<double>(chol, mu, sigma, dim, &host);
    const int rc = dev_alloc(&dev, host.size() * sizeof(double));
    if (rc != TQF_OK) return rc;
    e = cudaMemcpy(dev, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice);
  } else {
    std::vector<float> host, mma;
    build_split<float>(chol, mu, sigma, dim, &host);
    build_mma(chol, mu, sigma, dim, &mma);   // tensor-core kernels' tables follow
    host.insert(host.end(), mma.begin(), mma.end());
    build_tc5(chol, mu, sigma, dim, &mma);
    host.insert(host.end(), mma.begin(), mma.end());
    const int rc = dev_alloc(&dev, host.size() * sizeof(float));
    if (rc != TQF_OK) return rc;
    e = cudaMemcpy(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    dev_release(&dev, 1);
    return cuda_fail(e, "mvgbm_upload_split");
  }
  *out_dev = dev;
  return TQF_OK;
}

template <typename Real>
struct MvSplitCfg {
  static constexpr int NP = sizeof(Real) == 4 ? 2 : 1;       // paths per thread
  static constexpr int kPaths = kMvLanes * NP;                // paths per CTA chunk
  static constexpr int kIdxBits = kMvLowBits + (NP == 2 ? 1 : 0);
  static constexpr int kMinBlocks = sizeof(Real) == 4 ? 4 : 3;
  static constexpr size_t kDynSmem =
      (2 * kMvDim * kPaths + kMvParts * kMvSplitStride) * sizeof(Real);
};

template <typename Real>
__global__ void __launch_bounds__(kMvParts * kMvLanes, MvSplitCfg<Real>::kMinBlocks)
mvgbm_split_kernel(const __grid_constant__ MvParams<Real, kMvDim> P) {
  using Cfg = MvSplitCfg<Real>;
  constexpr int NP = Cfg::NP, kPaths = Cfg::kPaths;
  __shared__ uint32_t s_high[kMvTileDims];
  __shared__ uint4 s_low[kMvTileDims * 2];
  __shared__ double s_acc[TQF_MAX_PAYOFFS * 3];
  __shared__ Real s_red[kMvParts][kPaths];
  // dynamic: normals [2][16 groups][NP][32 lanes][4] | factor / mu / sigma
  extern __shared__ __align__(16) unsigned char s_mv_dyn[];
  Real* s_zall = reinterpret_cast<Real*>(s_mv_dyn);
  Real* s_split = s_zall + 2 * kMvDim * kPaths;
  const int tid = threadIdx.x, lane = tid & 31, part = tid >> 5;
  for (int i = tid; i < kMvParts * kMvSplitStride; i += blockDim.x) s_split[i] = P.lsplit[i];
  const fm::ConstTab tab(P.logtab);
  // the Philox stream's Box-Muller logarithm reads the MID-free table stored behind it
  const fm::ConstTab tab0(P.logtab + 2 * TQF_LOGTAB_COUNT);
  for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();

  // Sobol index of path q of this thread: chunk * kPaths + q * 32 + lane -> the
  // 5 low bits come from the lane, bit 5 (NP == 2) is q, the rest from the chunk
  uint32_t lowmask[kMvLowBits];
#pragma unroll
  for (int b = 0; b < kMvLowBits; ++b) {
    lowmask[b] = 0u - ((static_cast<uint32_t>(lane) >> b) & 1u);
    asm volatile("" : "+r"(lowmask[b]));
  }
  const int dim = P.dim;
  const int tile_steps = dim <= kMvTileDims ? kMvTileDims / dim : 1;
  const uint64_t stream_stride = static_cast<uint64_t>(P.num_steps_total) * dim;
  const uint64_t chunk_base = P.first_index & ~static_cast<uint64_t>(kPaths - 1);
  const uint64_t num_chunks = (P.first_index + P.path_count - chunk_base + kPaths - 1) / kPaths;
  const int j_begin = part * kMvRows;                                   // draws of this thread
  const int j_count = max(0, min(kMvRows, dim - j_begin));
  // shared-window addresses used by the step loop
  const uint32_t zall_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_zall));
  const uint32_t lp_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_split)) +
                           part * kMvSplitStride * sizeof(Real);
  const uint32_t low_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_low));
  const uint32_t high_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_high));
  // normal j of path q of this lane sits at (((j / 4) * NP + q) * 32 + lane) * 4 + j % 4;
  // this thread writes j = 16 part + jj: groups 4 part + jj / 4
  const uint32_t zlane_off = lane * 4 * sizeof(Real);
  const uint32_t zwrite_off = (part * 4 * NP * kMvLanes * 4) * sizeof(Real) + zlane_off;
  constexpr uint32_t kZq = kMvLanes * 4 * sizeof(Real);        // bytes between paths q, q + 1
  constexpr uint32_t kZg = NP * kZq;                            // bytes between groups

  for (uint64_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    bool valid[NP];
    uint64_t local[NP], elem_base[NP];
    Real x[NP][kMvRows];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const uint64_t index = chunk_base + chunk * kPaths + q * kMvLanes + lane;
      valid[q] = index >= P.first_index && index < P.first_index + P.path_count;
      local[q] = index - P.first_index;
      elem_base[q] = valid[q] ? (P.path_offset + local[q]) * stream_stride : 0;
#pragma unroll
      for (int r = 0; r < kMvRows; ++r) x[q][r] = P.x0[4 * r + part];
    }

    // payoffs / stores of the state after `step_index` steps
    auto record = [&](int step_index, int slot) {
      if (P.mode != MODE_PRICE) {
#pragma unroll
        for (int q = 0; q < NP; ++q)
          if (valid[q]) {
#pragma unroll
            for (int r = 0; r < kMvRows; ++r) {
              const int i = 4 * r + part;
              if (i < dim)
                P.out[static_cast<int64_t>(local[q]) * P.stride_path + slot * P.stride_time +
                      i * P.stride_dim] = P.store_exp ? static_cast<Real>(exp(x[q][r])) : x[q][r];
            }
          }
        return;
      }
      for (int pq = 0; pq < P.num_payoffs; ++pq) {
        const PayoffK& d = P.pay[pq];
        if (d.step != step_index) continue;
        // basket mean (component < 0) or one component, assembled through smem
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          Real part_val = 0;
#pragma unroll
          for (int r = 0; r < kMvRows; ++r) {
            const int i = 4 * r + part;
            const bool take = d.component < 0 ? i < dim : i == d.component;
            part_val += selp_real(x[q][r], Real(0), take ? 1 : 0);
          }
          s_red[part][q * kMvLanes + lane] = part_val;
        }
        __syncthreads();
        if (part == 0) {
          double sum = 0.0, sq = 0.0, bad = 0.0;
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const int c = q * kMvLanes + lane;
            Real m = (s_red[0][c] + s_red[1][c]) + (s_red[2][c] + s_red[3][c]);
            if (d.component < 0) m = m / static_cast<Real>(dim);
            if (valid[q]) {
              const double v = eval_payoff(d, static_cast<double>(m), 0.0, 0.0);
              if (isfinite(v)) {
                sum += v;
                sq += v * v;
              } else {
                bad += 1.0;
              }
            }
          }
          sum = warp_sum(sum);
          sq = warp_sum(sq);
          bad = warp_sum(bad);
          if (lane == 0) {
            s_acc[pq * 3 + 0] += sum;
            s_acc[pq * 3 + 1] += sq;
            s_acc[pq * 3 + 2] += bad;
          }
        }
      }
    };
    if (P.record_slot[0] >= 0) record(0, P.record_slot[0]);

    for (int s0 = 0; s0 < P.num_steps; s0 += tile_steps) {
      const int s1 = min(P.num_steps, s0 + tile_steps);
      if (P.rngk == RNGK_SOBOL) {
        __syncthreads();
        const uint32_t high_bits =
            static_cast<uint32_t>((chunk_base + chunk * kPaths) >> Cfg::kIdxBits);
        for (int dd = tid; dd < (s1 - s0) * dim; dd += blockDim.x) {
          // all 32 direction words at once (8 independent 16-byte loads), then the
          // XOR of the words selected by the chunk's high index bits
          const uint4* v4 = reinterpret_cast<const uint4*>(
              P.sobol_v + (static_cast<size_t>(s0) * dim + dd) * 32);
          uint4 w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) w[q] = __ldg(v4 + q);
          s_low[2 * dd] = w[0];
          s_low[2 * dd + 1] = w[1];
          const uint32_t* wv = reinterpret_cast<const uint32_t*>(w);
          uint32_t h = 0;
#pragma unroll
          for (int b = Cfg::kIdxBits; b < 32; ++b)
            h ^= wv[b] & (0u - ((high_bits >> (b - Cfg::kIdxBits)) & 1u));
          s_high[dd] = h;
        }
        __syncthreads();
      }
      for (int s = s0; s < s1; ++s) {
        const uint32_t zbuf_addr = zall_addr + (s & 1) * (kMvDim * kPaths) * sizeof(Real);
        const uint32_t zw = zbuf_addr + zwrite_off;
        // ---- this thread's 16 x NP normals of step s
        if (P.rngk == RNGK_SOBOL) {
          const uint32_t la = low_addr + ((s - s0) * dim + j_begin) * 32;
          const uint32_t ha = high_addr + ((s - s0) * dim + j_begin) * 4;
#pragma unroll 4
          for (int jj = 0; jj < kMvRows; ++jj) {
            if (jj < j_count) {
              const uint4 l0 = lds_u4(la + jj * 32);
              const uint4 l1 = lds_u4(la + jj * 32 + 16);
              uint32_t xb[NP];
              xb[0] = lds_u32(ha + jj * 4);
              xb[0] ^= l0.x & lowmask[0];
              xb[0] ^= l0.y & lowmask[1];
              xb[0] ^= l0.z & lowmask[2];
              xb[0] ^= l0.w & lowmask[3];
              xb[0] ^= l1.x & lowmask[4];
              if (NP == 2) xb[NP - 1] = xb[0] ^ l1.y;      // index bit 5 set
              Real zo[NP];
              sobol_normals<NP>(tab, xb, zo, P.sobol_clamp);
#pragma unroll
              for (int q = 0; q < NP; ++q)
                sts_real(zw + (jj >> 2) * kZg + q * kZq + (jj & 3) * sizeof(Real), zo[q]);
            } else {
#pragma unroll
              for (int q = 0; q < NP; ++q)
                sts_real(zw + (jj >> 2) * kZg + q * kZq + (jj & 3) * sizeof(Real), Real(0));
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            PhiloxStreamV<Real, 1> stream;
            const uint64_t fe[1] = {elem_base[q] + static_cast<uint64_t>(s) * dim + j_begin};
            if (j_count > 0) stream.init(P.key, P.ctr, tab0, fe);
            for (int jj = 0; jj < kMvRows; ++jj) {
              Real zo[1] = {0};
              if (jj < j_count) stream.next(P.key, P.ctr, tab0, zo);
              sts_real(zw + (jj >> 2) * kZg + q * kZq + (jj & 3) * sizeof(Real), zo[0]);
            }
          }
        }
        __syncthreads();   // the only barrier of the step (normals are double-buffered)
        const Real dt = P.coef[2 * s], sq = P.coef[2 * s + 1];
        mv_rows<Real, NP>(lp_addr, P.exact_log, zbuf_addr + zlane_off, dt, sq, x);
        const int flag = P.record_slot[s + 1];
        if (flag >= 0) record(s + 1, flag);
      }
    }
  }
  if (P.mode == MODE_PRICE) {
    __syncthreads();
    for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += blockDim.x) {
      const int q = i / 3, k = i - q * 3;
      P.partials[(static_cast<size_t>(blockIdx.x) * TQF_MAX_PAYOFFS + q) * 4 + k] = s_acc[i];
    }
  }
}

// ---------------------------------------------------------------------------
// float32, Sobol, 8 < dim <= 64: the triangular mat-vec of 16 paths at a time
// IS a dense contraction, [64 x 64] factor times [64 x 16] normals, and runs on
// the tensor cores (mma.sync m16n8k8, TF32 operands, FP32 accumulate).  A plain
// TF32 product (10-bit mantissa) would break the 1e-5 parity bound, so both
// operands are split, L = Lh + Ll (on the host), z = zh + zl (two instructions
// per draw), and Lh zh + Ll zh + Lh zl is accumulated: error ~2^-21 per term.
//   * a warp owns 16 paths (two n-tiles); the CTA's 4 warps = 64 Sobol indices;
//   * every thread draws its normals DIRECTLY in the B-fragment layout (lane
//     c = lane % 4 owns dimensions 8 kt + c and 8 kt + 4 + c of path lane / 4 of
//     each n-tile): no shared-memory staging of the normals and no barrier in
//     the step loop;
//   * the 20 non-zero 16 x 8 tiles of the factor sit in shared memory in
//     A-fragment order (one conflict-free LDS.128 per tile and split part);
//   * the state lives in the C-fragment layout (rows 16 mt + g, 16 mt + 8 + g,
//     paths 2 c, 2 c + 1 of each n-tile), where the Euler update is applied.
// Per path-step this issues ~2.2x fewer instructions than mvgbm_split_kernel
// (the 2 080 FFMAs become 7.5 warp-wide HMMAs).
constexpr int kMmaWarps = 4, kMmaPathsPerWarp = 16, kMmaPaths = kMmaWarps * kMmaPathsPerWarp;
constexpr int kMmaIdxBits = 6;
constexpr int kMmaTiles = 20;                                    // (mt, kt) with kt <= 2 mt + 1
constexpr int kMmaFragWords = kMmaTiles * 2 * 32 * 4;           // hi / lo fragments
constexpr int kMmaTabWords = kMmaFragWords + 2 * kMvDim;        // + mu[64], sigma[64]
#ifndef TQF_MMA_TILE_DIMS
#define TQF_MMA_TILE_DIMS 256
#endif
constexpr int kMmaTileDims = TQF_MMA_TILE_DIMS;   // Sobol dimensions staged at once (8 steps of 64; 256: +0.8 % time)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// Round to nearest TF32: the remainder z - zh then has a sign independent of
// z, so the tensor core's truncation of it does not bias |z| (a truncating
// split measurably lowers the C4 price by 4e-7 relative).
__device__ __forceinline__ uint32_t tf32_rna(float v) {
  // (cvt.rna.tf32.f32 is emulated with ~5 instructions on sm_100a)
  return (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
}

static uint32_t tf32_round_bits(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  b += 0x00000fffu + ((b >> 13) & 1u);
  return b & 0xffffe000u;
}

// Host: factor -> [tile][hi | lo][lane][4] A fragments (tile order: kt outer,
// mt = kt / 2 .. 3 inner -- the order the kernel consumes them), then mu, sigma.
static void build_mma(const double* chol, const double* mu, const double* sigma, int dim,
                      std::vector<float>* out) {
  out->assign(kMmaTabWords, 0.0f);
  uint32_t* w = reinterpret_cast<uint32_t*>(out->data());
  int ti = 0;
  for (int kt = 0; kt < 8; ++kt)
    for (int mt = kt / 2; mt < 4; ++mt, ++ti)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, c = lane & 3;
        const int rows[4] = {16 * mt + g, 16 * mt + g + 8, 16 * mt + g, 16 * mt + g + 8};
        const int cols[4] = {8 * kt + c, 8 * kt + c, 8 * kt + c + 4, 8 * kt + c + 4};
        for (int e = 0; e < 4; ++e) {
          const int i = rows[e], j = cols[e];
          // row i of the factor scaled by sigma_i: the kernel applies
          // x_i' = x_i (1 + mu_i dt + sqrt_dt sum_j (sigma_i L_ij) z_j)
          const float v = (j <= i && i < dim)
                              ? static_cast<float>(sigma[i] * chol[static_cast<size_t>(i) * dim + j])
                              : 0.0f;
          const uint32_t hb = tf32_round_bits(v);
          float hf;
          std::memcpy(&hf, &hb, 4);
          w[((ti * 2 + 0) * 32 + lane) * 4 + e] = hb;
          w[((ti * 2 + 1) * 32 + lane) * 4 + e] = tf32_round_bits(v - hf);
        }
      }
  for (int i = 0; i < dim; ++i) {
    (*out)[kMmaFragWords + i] = static_cast<float>(mu[i]);
    (*out)[kMmaFragWords + kMvDim + i] = static_cast<float>(sigma[i]);
  }
}

// float32 inverse normal CDF by table (tools/fit_ndtri_f32_table.py):
//   z = t f(a), a = 1 - t^2; the top 11 bits of a (8 exponent + 3 mantissa bits) select a
//   cubic in r = a - centre: 14 instructions per draw against 27 for the logarithm +
//   degree-10 polynomial, no tail branch; max relative error 2e-7 (mean 4e-8).
// The table lives in shared memory as EIGHT interleaved copies, copy k in the 16-byte
// bank group k: lane l reads copy l % 8, so the eight lanes that a 128-bit shared load
// serves per cycle never collide, whatever rows they ask for (a single copy would
// serialise ~2.7x on random rows and make the LSU the bottleneck).
constexpr int kNdTabRows = 186, kNdTabBase = 831;
constexpr int kNdTabSmemBytes = kNdTabRows * 8 * 16;
template <int N>
__device__ __forceinline__ void sobol_normals_f32_tab(const uint32_t (&xb)[N], float (&z)[N],
                                                      uint32_t tab_lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float t = fmaf(__uint2float_rn(xb[k]), 4.656612873077393e-10f, -1.0f);
    const float a = fmaf(-t, t, 1.0f);
    const uint32_t bits = __float_as_uint(a);
    const uint32_t idx = max(bits >> 20, static_cast<uint32_t>(kNdTabBase));
    const uint4 row = lds_u4(tab_lane + idx * 128u);          // tab_lane is pre-offset by -base
    const float r = a - __uint_as_float((bits & 0xFFF00000u) | 0x00080000u);
    const float q = fmaf(fmaf(fmaf(__uint_as_float(row.w), r, __uint_as_float(row.z)), r,
                              __uint_as_float(row.y)), r, __uint_as_float(row.x));
    z[k] = t * q;
  }
}

// FULL: dim == 64 (no padding checks in the draw loop).  TAB: table-driven draws.
// NS: steps whose accumulators are formed together (they share the A-fragment loads).
template <bool FULL, bool TAB, int NS>
__global__ void __launch_bounds__(kMmaWarps * 32, NS == 1 ? 4 : 3)
mvgbm_mma_kernel(const __grid_constant__ MvParams<float, kMvDim> P) {
  extern __shared__ __align__(16) unsigned char s_ndt[];   // TAB: [rows][8 copies][4 floats]
  // per staged dimension 8 words: hi ^ (warp bits) for the 4 warps, then the
  // direction words of index bits 0..3
  __shared__ uint4 s_sob[kMmaTileDims * 2];
  __shared__ uint4 s_frag[kMmaFragWords / 4];
  __shared__ double s_acc[kMmaWarps][TQF_MAX_PAYOFFS * 3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const float* tabp = P.lsplit + kMvParts * kMvSplitStride;
  for (int i = tid; i < kMmaFragWords / 4; i += blockDim.x)
    s_frag[i] = __ldg(reinterpret_cast<const uint4*>(tabp) + i);
  for (int i = tid; i < kMmaWarps * TQF_MAX_PAYOFFS * 3; i += blockDim.x) (&s_acc[0][0])[i] = 0.0;
  if (TAB) {
    uint4* dst = reinterpret_cast<uint4*>(s_ndt);
    for (int i = tid; i < kNdTabRows * 8; i += blockDim.x) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(P.ndtab) + (i >> 3));
      // row 0 = the draws with t = +-1 (float32 uniform rounded to 0 or 1): +-inf
      // like the reference, or +-ndtri(1 - 2^-24) in the clamped mode
      if ((i >> 3) == 0 && P.sobol_clamp) v.x = __float_as_uint(5.4199314f);
      dst[i] = v;
    }
  }
  const uint32_t tab_lane = static_cast<uint32_t>(__cvta_generic_to_shared(s_ndt)) + (lane & 7) * 16 -
                            kNdTabBase * 128u;
  // rows of this thread in the C layout: 16 mt + g (h = 0), 16 mt + 8 + g (h = 1)
  float mu[8], x0[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = 16 * (r >> 1) + 8 * (r & 1) + g;
    mu[r] = __ldg(tabp + kMmaFragWords + i);
    x0[r] = P.x0[i];
  }
  __syncthreads();

  // index of the path this thread DRAWS for, n-tile 0: warp * 16 + g (bit 3 = n-tile)
  uint32_t lowmask[3];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    lowmask[b] = 0u - ((static_cast<uint32_t>(g) >> b) & 1u);
    asm volatile("" : "+r"(lowmask[b]));
  }
  const int dim = P.dim;
  const int tile_steps = kMmaTileDims / dim;   // dim <= 64 -> >= 4
  const uint64_t chunk_base = P.first_index & ~static_cast<uint64_t>(kMmaPaths - 1);
  const uint64_t num_chunks =
      (P.first_index + P.path_count - chunk_base + kMmaPaths - 1) / kMmaPaths;
  const uint32_t sob_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_sob));
  const uint32_t frag_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_frag)) + lane * 16;
  // this lane's first dimension: direction words at +16, the warp's high word at + 4 warp
  const uint32_t sob_v = sob_addr + c * 32 + 16, sob_h = sob_addr + c * 32 + warp * 4;

  for (uint64_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    // paths of this thread in the C layout: warp * 16 + nt * 8 + 2 c + e
    bool valid[2][2];
    uint64_t local[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint64_t index = chunk_base + chunk * kMmaPaths + warp * 16 + nt * 8 + 2 * c + e;
        valid[nt][e] = index >= P.first_index && index < P.first_index + P.path_count;
        local[nt][e] = index - P.first_index;
      }
    float x[4][2][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) x[mt][nt][e] = x0[mt * 2 + (e >> 1)];

    auto record = [&](int step_index, int slot) {
      if (P.mode != MODE_PRICE) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 16 * mt + 8 * (e >> 1) + g;
              if (valid[nt][e & 1] && i < dim) {
                const float v = x[mt][nt][e];
                P.out[static_cast<int64_t>(local[nt][e & 1]) * P.stride_path + slot * P.stride_time +
                      i * P.stride_dim] = P.store_exp ? expf(v) : v;
              }
            }
        return;
      }
      for (int pq = 0; pq < P.num_payoffs; ++pq) {
        const PayoffK& d = P.pay[pq];
        if (d.step != step_index) continue;
        // basket sum (component < 0) or one component: own rows, then the 8 lanes
        // that share c (shuffles over the g bits)
        float part[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float sacc = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 16 * mt + 8 * h + g;
                const bool take = d.component < 0 ? i < dim : i == d.component;
                sacc += take ? x[mt][nt][2 * h + e] : 0.0f;
              }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) sacc += __shfl_xor_sync(0xFFFFFFFFu, sacc, o);
            part[nt][e] = sacc;
          }
        double sum = 0.0, sq = 0.0, bad = 0.0;
        if (g == 0) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float m = part[nt][e];
              if (d.component < 0) m = m / static_cast<float>(dim);
              if (valid[nt][e]) {
                const double v = eval_payoff(d, static_cast<double>(m), 0.0, 0.0);
                if (isfinite(v)) {
                  sum += v;
                  sq += v * v;
                } else {
                  bad += 1.0;
                }
              }
            }
        }
        sum = warp_sum(sum);
        sq = warp_sum(sq);
        bad = warp_sum(bad);
        if (lane == 0) {
          s_acc[warp][pq * 3 + 0] += sum;
          s_acc[warp][pq * 3 + 1] += sq;
          s_acc[warp][pq * 3 + 2] += bad;
        }
      }
    };
    if (P.record_slot[0] >= 0) record(0, P.record_slot[0]);

    for (int s0 = 0; s0 < P.num_steps; s0 += tile_steps) {
      const int s1 = min(P.num_steps, s0 + tile_steps);
      __syncthreads();
      {
        const uint32_t high_bits =
            static_cast<uint32_t>((chunk_base + chunk * kMmaPaths) >> kMmaIdxBits);
        // Eight lanes per dimension, one 16-byte quad of its 32 direction words each:
        // consecutive lanes read consecutive 16 bytes (fully coalesced; one thread per
        // dimension reading 128 bytes cost 8x the LSU wavefronts, a quarter of the
        // kernel's LSU budget).  The XOR over the chunk's high index bits is combined
        // with three shuffles.
        const int q = tid & 7;
        const int ndims = (s1 - s0) * dim;
        for (int base = 0; base < ndims; base += blockDim.x / 8) {
          const int dd = base + (tid >> 3);
          const bool live_dd = dd < ndims;
          uint4 w = make_uint4(0u, 0u, 0u, 0u);
          if (live_dd)
            w = __ldg(reinterpret_cast<const uint4*>(
                          P.sobol_v + (static_cast<size_t>(s0) * dim + dd) * 32) + q);
          const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
          uint32_t h = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = 4 * q + j;                 // index bit of this word
            const uint32_t bit = b >= kMmaIdxBits ? (high_bits >> (b - kMmaIdxBits)) & 1u : 0u;
            h ^= wv[j] & (0u - bit);
          }
          h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 1);
          h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 2);
          h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 4);
          if (live_dd && q == 1) s_sob[2 * dd] = make_uint4(h, h ^ w.x, h ^ w.y, h ^ w.x ^ w.y);
          if (live_dd && q == 0) s_sob[2 * dd + 1] = w;
        }
      }
      __syncthreads();
      // NS consecutive steps share every load of the factor's A fragments: sigma L z is
      // state-independent, so the accumulators of step s + 1 can be formed next to those
      // of step s before either update runs.  (The A fragments are 41% of the kernel's
      // shared-memory wavefronts, and the LSU data pipe is what bounds it: ncu
      // l1tex__data_pipe_lsu_wavefronts 91%.)
      for (int s = s0; s < s1; s += NS) {
        float acc[NS][4][2][4];
#pragma unroll
        for (int u = 0; u < NS; ++u)
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[u][mt][nt][e] = 0.0f;
        uint32_t soff[NS];
#pragma unroll
        for (int u = 0; u < NS; ++u)        // a step beyond the tile re-reads the last one (discarded)
          soff[u] = (min(s + u, s1 - 1) - s0) * dim * 32;
        // (A rolled loop over k-tile pairs with the row tiles predicated -- 9 KB of
        // code instead of 21 KB -- measured 767 ms against 730 ms: the fetch stalls
        // it removes cost less than the scheduling freedom it takes away.)
        int ti = 0;   // compile-time after unrolling
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
          // this thread's entries of the B fragments of k-tile kt: dimensions
          // 8 kt + c (h = 0) and 8 kt + 4 + c (h = 1) of both n-tiles, drawn side by side
          uint32_t bh[NS][2][2], bl[NS][2][2];
#pragma unroll
          for (int u = 0; u < NS; ++u) {
            uint32_t xb[4];
            bool live[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = 8 * kt + 4 * h;           // + c
              live[h] = FULL || j + c < dim;
              const uint32_t jo = (FULL || live[h]) ? j * 32 : 0;   // stay inside the staged tile
              const uint4 v = lds_u4(sob_v + soff[u] + jo);
              uint32_t x = lds_u32(sob_h + soff[u] + jo);
              x ^= v.x & lowmask[0];
              x ^= v.y & lowmask[1];
              x ^= v.z & lowmask[2];
              xb[2 * h] = x;
              xb[2 * h + 1] = x ^ v.w;                 // index bit 3 = n-tile
            }
            float z[4];
            if (TAB)
              sobol_normals_f32_tab<4>(xb, z, tab_lane);
            else
              sobol_normals_f32<4>(xb, z, P.sobol_clamp);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const float zz = (FULL || live[h]) ? z[2 * h + nt] : 0.0f;
                bh[u][nt][h] = tf32_rna(zz);
                bl[u][nt][h] = __float_as_uint(zz - __uint_as_float(bh[u][nt][h]));
              }
          }
#pragma unroll
          for (int mt = kt / 2; mt < 4; ++mt) {
            const uint4 ah = lds_u4(frag_addr + ((ti + mt - kt / 2) * 2 + 0) * 512);
            const uint4 al = lds_u4(frag_addr + ((ti + mt - kt / 2) * 2 + 1) * 512);
            // small terms first; the three products of one accumulator are spaced
            // by the other accumulators
#pragma unroll
            for (int u = 0; u < NS; ++u)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[u][mt][nt], al, bh[u][nt][0], bh[u][nt][1]);
#pragma unroll
            for (int u = 0; u < NS; ++u)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[u][mt][nt], ah, bl[u][nt][0], bl[u][nt][1]);
#pragma unroll
            for (int u = 0; u < NS; ++u)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[u][mt][nt], ah, bh[u][nt][0], bh[u][nt][1]);
          }
          ti += 4 - kt / 2;
        }
        // acc_i = sum_j sigma_i L_ij z_j.  Euler: x_i' = x_i + x_i (mu_i dt + sqrt_dt acc_i);
        // exact log-normal increment: x_i' = x_i + (mu_i dt + sqrt_dt acc_i), mu = means - vols^2 / 2.
        // (The relative increment is NOT merged into 1 + ...: rounding 1 + mu dt to
        // float32 is the same error for every path and step, a 1e-4 bias of the price.)
#pragma unroll
        for (int u = 0; u < NS; ++u) {
          if (s + u >= s1) break;
          const float dt = P.coef[2 * (s + u)], sq = P.coef[2 * (s + u) + 1];
          float c1[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) c1[r] = mu[r] * dt;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float f = fmaf(sq, acc[u][mt][nt][e], c1[mt * 2 + (e >> 1)]);
                x[mt][nt][e] = P.exact_log ? x[mt][nt][e] + f : fmaf(x[mt][nt][e], f, x[mt][nt][e]);
              }
          const int flag = P.record_slot[s + u + 1];
          if (flag >= 0) record(s + u + 1, flag);
        }
      }
    }
  }
  if (P.mode == MODE_PRICE) {
    __syncthreads();
    for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += blockDim.x) {
      const int q = i / 3, k = i - q * 3;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kMmaWarps; ++w) v += s_acc[w][i];
      P.partials[(static_cast<size_t>(blockIdx.x) * TQF_MAX_PAYOFFS + q) * 4 + k] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// float32, Sobol, dim == 64, fused price: the same contraction on the 5th-generation
// tensor cores (tcgen05.mma.kind::tf32, accumulator in tensor memory).
//
//   * a CTA owns a tile of 128 consecutive Sobol indices; path p IS TMEM lane p: it
//     holds row p of the A operand (the 64 scaled normals) and row p of the
//     accumulator (the 64 increments).  Two threads carry a path -- warps 0-3 its
//     factors / assets 0-31, warps 4-7 the other 32 (both reach lanes 32 (w % 4) ..)
//     -- with their half of the state in registers for all steps: no fragment
//     layouts, no shuffles, 8 warps per CTA for the same tensor memory;
//   * per step a thread draws its 32 normals (16 side by side), scales them by
//     sqrt(dt), splits them into TF32 hi / lo parts and writes them to TMEM with
//     tcgen05.st (32x32b.x16); column 64 of A carries dt (hi / lo), so that the
//     drift mu_i dt comes out of the same contraction (row 64 of B = mu);
//   * B = (diag(sigma) L)^T, hi and lo parts, sits in shared memory for the whole
//     kernel in the canonical K-major no-swizzle UMMA layout (8 x 16-byte core
//     matrices, LBO = 1024 B between K chunks, SBO = 128 B between row groups);
//   * one thread issues 27 MMAs (128 x 64 x 8 each): A_lo B_hi, A_hi B_lo, A_hi B_hi
//     (9 K slices each), tcgen05.commit -> mbarrier; every thread then reads its row of
//     the accumulator with tcgen05.ld and applies x += x f (or x += f);
//   * the Sobol integers are T[d][lane] ^ H[d][warp]: T = the 32 combinations of the
//     direction words of index bits 0-4, H = index bits 5-6 (the warp) and the tile's
//     high bits, both staged in shared memory one step ahead (double buffered).
// Two CTAs per SM (256 of the 512 TMEM columns each): while one waits for its
// MMAs the other draws.
constexpr int kT5Paths = 128;                         // paths per tile = TMEM lanes
constexpr int kT5Threads = 256;                       // two threads per path, 32 factors each
constexpr int kT5Half = kMvDim / 2;
constexpr int kT5K = 72;                              // 64 factors + dt column, padded to 8
constexpr int kT5PartBytes = (kT5K / 4) * 1024;       // one part (hi or lo) of B
constexpr int kT5BWords = 2 * kT5PartBytes / 4;
constexpr int kT5TmemCols = 256;
constexpr uint32_t kT5ColD = 0, kT5ColAh = 64, kT5ColAl = 64 + kT5K;
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1),
// A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kT5Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int kT5SmemB = 0;
constexpr int kT5SmemTab = kT5SmemB + 2 * kT5PartBytes;
constexpr int kT5SmemT = kT5SmemTab + kNdTabSmemBytes;             // 2 x [64][32] words
constexpr int kT5HStride = kMvDim + 8;                // +8: the four warp variants of a dimension land in different banks
constexpr int kT5SmemH = kT5SmemT + 2 * 64 * 32 * 4;               // 2 x [4][72] words
constexpr int kT5SmemAcc = kT5SmemH + 2 * 4 * kT5HStride * 4;              // [4][MAX_PAYOFFS * 3] doubles
constexpr int kT5SmemPart = kT5SmemAcc + 4 * TQF_MAX_PAYOFFS * 3 * 8;   // [128] floats
constexpr int kT5SmemBar = kT5SmemPart + kT5Paths * 4;
constexpr int kT5SmemBytes = kT5SmemBar + 16;

// Host: B[k][n] = sigma_n L_nk (k < 64), mu_n (k = 64), 0 beyond; element (n, k) of a part at
// float index (k / 4) * 256 + n * 4 + k % 4.
static void build_tc5(const double* chol, const double* mu, const double* sigma, int dim,
                      std::vector<float>* out) {
  out->assign(kT5BWords, 0.0f);
  uint32_t* w = reinterpret_cast<uint32_t*>(out->data());
  for (int n = 0; n < kMvDim; ++n)
    for (int k = 0; k < kT5K; ++k) {
      float v = 0.0f;
      if (n < dim && k < dim && k <= n) v = static_cast<float>(sigma[n] * chol[static_cast<size_t>(n) * dim + k]);
      if (n < dim && k == kMvDim) v = static_cast<float>(mu[n]);
      const uint32_t hb = tf32_round_bits(v);
      float hf;
      std::memcpy(&hf, &hb, 4);
      const int at = (k / 4) * 256 + n * 4 + (k % 4);
      w[at] = hb;
      w[kT5PartBytes / 4 + at] = tf32_round_bits(v - hf);
    }
}

__device__ __forceinline__ void t5_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One term (A part at a_col, B part at byte offset b_off) of a step: K slices 0..7 of the
// factors and slice 8 = the dt column.  The factor is lower triangular: slice j (factors
// 8 j .. 8 j + 7) only moves assets n >= 8 j, so it runs on the columns 16 (j / 2) .. 63 of
// the accumulator (N = 64, 64, 48, 48, 32, 32, 16, 16: 5/8 of the tensor-pipe time; N must be
// a multiple of 16 at M = 128).  `fresh`: the first slice overwrites the accumulator.
__device__ __forceinline__ void t5_term(uint32_t tmem, uint32_t a_col, uint64_t b_desc, uint32_t b_off,
                                        bool fresh) {
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int n0 = j < 8 ? 16 * (j / 2) : 0;
    const uint32_t idesc = (kT5Idesc & ~(0x3Fu << 17)) | (static_cast<uint32_t>((kMvDim - n0) >> 3) << 17);
    t5_mma(tmem + kT5ColD + n0, tmem + a_col + 8 * j, b_desc + ((b_off + j * 2048 + n0 * 16) >> 4), idesc,
           (fresh && j == 0) ? 0u : 1u);
  }
}

__device__ __forceinline__ void t5_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void t5_st8(uint32_t taddr, uint32_t v0) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %2, %2, %2, %2, %2, %2};\n"
               :
               : "r"(taddr), "r"(v0), "r"(z)
               : "memory");
}

__device__ __forceinline__ void t5_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Bounded wait: a protocol error traps (the launch fails) instead of hanging the device.
__device__ __forceinline__ void t5_mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// Stages step s of the tile whose paths have index >> 7 == high_bits: sT[d][l] = XOR of the
// direction words of the set bits of l (index bits 0-4), sH[w][d] = index bits 5-6 = w and the
// high bits.  Eight lanes per dimension read its 32 direction words as one coalesced 128 bytes
// (t5_stage_load, issued early: the words come from L2) and combine them with shuffles.
constexpr int kT5StageIters = kMvDim * 8 / kT5Threads;
struct T5Words {
  uint4 w[kT5StageIters];
};
// (dimensions beyond `dim` of a narrower model are staged as zero words: their draws are
// discarded by the kernel)
template <bool FULL>
__device__ __forceinline__ T5Words t5_stage_load(const uint32_t* __restrict__ sobol_v, int s, int dim_,
                                                 int tid) {
  const int dim = FULL ? kMvDim : dim_;
  T5Words r;
#pragma unroll
  for (int it = 0; it < kT5StageIters; ++it) {
    const int dd = it * (kT5Threads / 8) + (tid >> 3);
    r.w[it] = make_uint4(0u, 0u, 0u, 0u);
    if (FULL || dd < dim)
      r.w[it] = __ldg(reinterpret_cast<const uint4*>(sobol_v + (static_cast<size_t>(s) * dim + dd) * 32) +
                      (tid & 7));
  }
  return r;
}
__device__ __forceinline__ void t5_stage_write(const T5Words& r, uint32_t high_bits, uint32_t* sT,
                                               uint32_t* sH, int tid) {
  const int q = tid & 7, lane = tid & 31, base_lane = lane & ~7;
#pragma unroll
  for (int it = 0; it < kT5StageIters; ++it) {
    const int dd = it * (kT5Threads / 8) + (tid >> 3);
    const uint4 w = r.w[it];
    const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
    uint32_t h = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = 4 * q + j;                 // index bit of this word
      const uint32_t bit = b >= 7 ? (high_bits >> (b - 7)) & 1u : 0u;
      h ^= wv[j] & (0u - bit);
    }
    h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 1);
    h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 2);
    h ^= __shfl_xor_sync(0xFFFFFFFFu, h, 4);
    const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, w.x, base_lane);
    const uint32_t v1 = __shfl_sync(0xFFFFFFFFu, w.y, base_lane);
    const uint32_t v2 = __shfl_sync(0xFFFFFFFFu, w.z, base_lane);
    const uint32_t v3 = __shfl_sync(0xFFFFFFFFu, w.w, base_lane);
    const uint32_t v4 = __shfl_sync(0xFFFFFFFFu, w.x, base_lane + 1);
    const uint32_t v5 = __shfl_sync(0xFFFFFFFFu, w.y, base_lane + 1);
    const uint32_t v6 = __shfl_sync(0xFFFFFFFFu, w.z, base_lane + 1);
    // entries l = 4 q + {0, 1, 2, 3}: bits 2, 3, 4 of l are the bits of q
    const uint32_t e = (v2 & (0u - (q & 1u))) ^ (v3 & (0u - ((q >> 1) & 1u))) ^ (v4 & (0u - ((q >> 2) & 1u)));
    *reinterpret_cast<uint4*>(sT + dd * 32 + 4 * q) = make_uint4(e, e ^ v0, e ^ v1, e ^ v0 ^ v1);
    if (q < 4) sH[q * kT5HStride + dd] = h ^ (v5 & (0u - (q & 1u))) ^ (v6 & (0u - ((q >> 1) & 1u)));
  }
}
template <bool FULL>
__device__ __forceinline__ void t5_stage(const uint32_t* __restrict__ sobol_v, int s, int dim,
                                         uint32_t high_bits, uint32_t* sT, uint32_t* sH, int tid) {
  t5_stage_write(t5_stage_load<FULL>(sobol_v, s, dim, tid), high_bits, sT, sH, tid);
}

// FULL: dim == 64 (no padding checks in the draw loop).  PRICE: fused payoff reduction
// (otherwise the recorded states are stored).
template <bool FULL, bool PRICE>
__global__ void __launch_bounds__(kT5Threads, 2)
mvgbm_tc5_kernel(const __grid_constant__ MvParams<float, kMvDim> P) {
  extern __shared__ __align__(1024) unsigned char t5_smem[];
  // thread = (path p of the tile, half h of its factors / assets): warps 0-3 are half 0,
  // warps 4-7 half 1; warp w reaches TMEM lanes 32 (w % 4) .. + 31 = its paths
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pw = warp & 3, half = warp >> 2;
  const int p = tid & (kT5Paths - 1);
  uint32_t* sB = reinterpret_cast<uint32_t*>(t5_smem + kT5SmemB);
  uint32_t* sT = reinterpret_cast<uint32_t*>(t5_smem + kT5SmemT);
  uint32_t* sH = reinterpret_cast<uint32_t*>(t5_smem + kT5SmemH);
  double* s_acc = reinterpret_cast<double*>(t5_smem + kT5SmemAcc);
  float* s_part = reinterpret_cast<float*>(t5_smem + kT5SmemPart);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t5_smem + kT5SmemBar + 8);
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(t5_smem));
  const uint32_t bar = smem_base + kT5SmemBar;

  // ---- one-time set-up: B, the inverse-CDF table, the barrier, tensor memory
  const float* tabp = P.lsplit + kMvParts * kMvSplitStride + kMmaTabWords;
  for (int i = tid; i < kT5BWords / 4; i += kT5Threads)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(tabp) + i);
  {
    uint4* dst = reinterpret_cast<uint4*>(t5_smem + kT5SmemTab);
    for (int i = tid; i < kNdTabRows * 8; i += kT5Threads) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(P.ndtab) + (i >> 3));
      if ((i >> 3) == 0 && P.sobol_clamp) v.x = __float_as_uint(5.4199314f);
      dst[i] = v;
    }
  }
  for (int i = tid; i < 4 * TQF_MAX_PAYOFFS * 3; i += kT5Threads) s_acc[i] = 0.0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :
                 : "r"(smem_base + kT5SmemBar + 8), "r"(kT5TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  // B was written through the generic proxy; the tensor core reads it through the async proxy
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *s_tmem;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(pw * 32) << 16);   // this warp's lanes
  const uint32_t tab_lane = smem_base + kT5SmemTab + (lane & 7) * 16 - kNdTabBase * 128u;
  // shared-memory matrix descriptor of B (cute::UMMA::SmemDescriptor): start >> 4,
  // LBO = 1024 B (K chunks) at bit 16, SBO = 128 B (8-row groups) at bit 32, version 1 at bit 46
  const uint64_t b_desc = static_cast<uint64_t>(((smem_base + kT5SmemB) & 0x3FFFFu) >> 4) |
                          (static_cast<uint64_t>(1024 >> 4) << 16) |
                          (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);

  const uint64_t chunk_base = P.first_index & ~static_cast<uint64_t>(kT5Paths - 1);
  const uint64_t num_chunks = (P.first_index + P.path_count - chunk_base + kT5Paths - 1) / kT5Paths;
  uint32_t phase = 0;

  for (uint64_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    const uint64_t tile_index = chunk_base + chunk * kT5Paths;
    const uint64_t index = tile_index + p;
    const bool valid = index >= P.first_index && index < P.first_index + P.path_count;
    const uint32_t high_bits = static_cast<uint32_t>(tile_index >> 7);
    float x[kT5Half];                       // assets 32 half .. 32 half + 31 of path p
#pragma unroll
    for (int i = 0; i < kT5Half; ++i) x[i] = P.x0[half * kT5Half + i];

    // (called by all threads together: d.step and the payoff list are uniform)
    auto record = [&](int step_index, int slot) {
      const int dim = P.dim;
      if (!PRICE) {
        // path materialisation: every thread stores its 32 assets of path p
        if (valid) {
          float* o = P.out + static_cast<int64_t>(index - P.first_index) * P.stride_path +
                     static_cast<int64_t>(slot) * P.stride_time;
#pragma unroll
          for (int i = 0; i < kT5Half; ++i) {
            const int a = half * kT5Half + i;
            if (FULL || a < dim) o[a * P.stride_dim] = P.store_exp ? expf(x[i]) : x[i];
          }
        }
        return;
      }
      for (int pq = 0; pq < P.num_payoffs; ++pq) {
        const PayoffK& d = P.pay[pq];
        if (d.step != step_index) continue;
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < kT5Half; ++i) {
          const int a = half * kT5Half + i;
          const bool take = d.component < 0 ? a < dim : a == d.component;
          m += take ? x[i] : 0.0f;
        }
        __syncthreads();
        if (half == 1) s_part[p] = m;
        __syncthreads();
        double sum = 0.0, sq = 0.0, bad = 0.0;
        if (half == 0) {
          m += s_part[p];
          if (d.component < 0) m = m / static_cast<float>(dim);
          if (valid) {
            const double v = eval_payoff(d, static_cast<double>(m), 0.0, 0.0);
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
            double* acc = s_acc + (pw * TQF_MAX_PAYOFFS + pq) * 3;
            acc[0] += sum;
            acc[1] += sq;
            acc[2] += bad;
          }
        }
      }
    };
    if (P.record_slot[0] >= 0) record(0, P.record_slot[0]);

    __syncthreads();                       // every reader of the staging buffers is done
    t5_stage<FULL>(P.sobol_v, 0, P.dim, high_bits, sT, sH, tid);
    __syncthreads();

#pragma unroll 1
    for (int s = 0; s < P.num_steps; ++s) {
      const int buf = s & 1;
      const uint32_t* sTb = sT + buf * (kMvDim * 32) + half * (kT5Half * 32) + lane;
      const uint32_t* sHb = sH + buf * (4 * kT5HStride) + pw * kT5HStride + half * kT5Half;
      const float dt = P.coef[2 * s], sqdt = P.coef[2 * s + 1];
      T5Words next;
      if (s + 1 < P.num_steps) next = t5_stage_load<FULL>(P.sobol_v, s + 1, P.dim, tid);
      // ---- the 32 scaled normals of this half of the path -> A (hi / lo) in tensor memory
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t xb[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const uint4 hw = *reinterpret_cast<const uint4*>(sHb + c * 16 + i4 * 4);
          xb[i4 * 4 + 0] = sTb[(c * 16 + i4 * 4 + 0) * 32] ^ hw.x;
          xb[i4 * 4 + 1] = sTb[(c * 16 + i4 * 4 + 1) * 32] ^ hw.y;
          xb[i4 * 4 + 2] = sTb[(c * 16 + i4 * 4 + 2) * 32] ^ hw.z;
          xb[i4 * 4 + 3] = sTb[(c * 16 + i4 * 4 + 3) * 32] ^ hw.w;
        }
        float z[16];
        sobol_normals_f32_tab<16>(xb, z, tab_lane);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // (a narrower model: the padded factors multiply zero rows of B, but inf * 0 is NaN)
          const float zs = (FULL || half * kT5Half + c * 16 + i < P.dim) ? z[i] * sqdt : 0.0f;
          hi[i] = tf32_rna(zs);
          lo[i] = __float_as_uint(zs - __uint_as_float(hi[i]));
        }
        t5_st16(t_lane + kT5ColAh + half * kT5Half + c * 16, hi);
        t5_st16(t_lane + kT5ColAl + half * kT5Half + c * 16, lo);
      }
      if (half == 1) {
        const uint32_t dh = tf32_rna(dt);
        t5_st8(t_lane + kT5ColAh + kMvDim, dh);
        t5_st8(t_lane + kT5ColAl + kMvDim, __float_as_uint(dt - __uint_as_float(dh)));
      }
      // ---- the tables of the next step, while this step's stores drain
      if (s + 1 < P.num_steps)
        t5_stage_write(next, high_bits, sT + (buf ^ 1) * (kMvDim * 32), sH + (buf ^ 1) * (4 * kT5HStride), tid);
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // small terms first: A_lo B_hi, A_hi B_lo, A_hi B_hi (one K = 8 slice = 2 chunks = 2 KB of B)
        t5_term(tmem, kT5ColAl, b_desc, 0, true);
        t5_term(tmem, kT5ColAh, b_desc, kT5PartBytes, false);
        t5_term(tmem, kT5ColAh, b_desc, 0, false);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                     :
                     : "r"(bar)
                     : "memory");
      }
      t5_mbar_wait(bar, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // ---- f_i = sqrt_dt sum_k sigma_i L_ik z_k + mu_i dt; Euler x += x f, exact log step x += f
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float f[16];
        t5_ld16(t_lane + kT5ColD + half * kT5Half + c * 16, f);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          x[c * 16 + i] = P.exact_log ? x[c * 16 + i] + f[i] : fmaf(x[c * 16 + i], f[i], x[c * 16 + i]);
      }
      if (P.record_slot[s + 1] >= 0) record(s + 1, P.record_slot[s + 1]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kT5TmemCols)
                 : "memory");
  }
  if (PRICE) {
    for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += kT5Threads) {
      const int q = i / 3, k = i - q * 3;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 4; ++w) v += s_acc[w * TQF_MAX_PAYOFFS * 3 + i];
      P.partials[(static_cast<size_t>(blockIdx.x) * TQF_MAX_PAYOFFS + q) * 4 + k] = v;
    }
  }
}

// TQF_MVGBM_TC5=0 falls back to the mma.sync kernel (kept for the A/B; float64 and
// non-Sobol draws take the split kernel).
static bool tc5_enabled() {
  const char* e = std::getenv("TQF_MVGBM_TC5");   // read per launch: tests toggle it
  return !(e && e[0] == '0');
}

static bool mma_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("TQF_MVGBM_MMA");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename Real, int DMAX>
static void launch_split(const MvParams<Real, DMAX>& P, int grid, cudaStream_t stream) {
  if constexpr (DMAX == kMvDim) {
    const size_t smem = MvSplitCfg<Real>::kDynSmem;
    if (smem > 32 * 1024)
      cudaFuncSetAttribute(mvgbm_split_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem));
    mvgbm_split_kernel<Real><<<grid, kMvParts * kMvLanes, smem, stream>>>(P);
  }
}

template <typename Real, int DMAX>
static int launch_mv(const MvLaunch& a, cudaStream_t stream, int* grid_out) {
  // large (up to 19 KB): on the heap, one per call -- no shared state between
  // threads / streams / devices of one process (the launch copies it)
  std::unique_ptr<MvParams<Real, DMAX>> holder(new (std::nothrow) MvParams<Real, DMAX>);
  TQF_REQUIRE(holder, "out of memory");
  MvParams<Real, DMAX>& P = *holder;
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
  P.lsplit = static_cast<const Real*>(a.lsplit_dev);
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
  P.sobol_clamp = a.sobol_clamp;
  for (int i = 0; i < DMAX; ++i) {
    const bool in = i < a.dim;
    P.x0[i] = in ? static_cast<Real>(a.x0[i]) : Real(0);
    P.mu[i] = in ? static_cast<Real>(a.mu[i]) : Real(0);
    P.sigma[i] = in ? static_cast<Real>(a.sigma[i]) : Real(0);
    for (int j = 0; j <= i; ++j)
      P.L[i * (i + 1) / 2 + j] =
          (in && j < a.dim) ? static_cast<Real>(a.chol[static_cast<size_t>(i) * a.dim + j]) : Real(0);
  }
  if (DMAX == kMvDim) {
    constexpr uint64_t kPaths = MvSplitCfg<Real>::kPaths;
    const uint64_t base32 = a.first_index & ~(kPaths - 1);
    const uint64_t chunks32 = (a.first_index + a.path_count - base32 + kPaths - 1) / kPaths;
    int grid = static_cast<int>(chunks32 < static_cast<uint64_t>(a.max_grid)
                                    ? chunks32 : static_cast<uint64_t>(a.max_grid));
    if (grid < 1) grid = 1;
    *grid_out = grid;
    if constexpr (sizeof(Real) == 4 && DMAX == kMvDim) {
      if (a.rngk == RNGK_SOBOL && a.ndtab != nullptr && tc5_enabled()) {
        P.ndtab = a.ndtab;
        const uint64_t base128 = a.first_index & ~static_cast<uint64_t>(kT5Paths - 1);
        const uint64_t chunks128 = (a.first_index + a.path_count - base128 + kT5Paths - 1) / kT5Paths;
        int g5 = static_cast<int>(chunks128 < static_cast<uint64_t>(2 * kSMs) ? chunks128 : 2 * kSMs);
        if (g5 < 1) g5 = 1;
        if (g5 > a.max_grid) g5 = a.max_grid;      // partials hold max_grid rows
        *grid_out = g5;
        const bool price = a.mode == MODE_PRICE;
        auto k5 = a.dim == kMvDim ? (price ? mvgbm_tc5_kernel<true, true> : mvgbm_tc5_kernel<true, false>)
                                  : (price ? mvgbm_tc5_kernel<false, true> : mvgbm_tc5_kernel<false, false>);
        TQF_CUDA_OK(cudaFuncSetAttribute(k5, cudaFuncAttributeMaxDynamicSharedMemorySize, kT5SmemBytes));
        k5<<<g5, kT5Threads, kT5SmemBytes, stream>>>(P);
        TQF_CUDA_OK(cudaGetLastError());
        return TQF_OK;
      }
      if (a.rngk == RNGK_SOBOL && mma_enabled()) {
        static const bool tab = [] {
          const char* e = std::getenv("TQF_MVGBM_NDTRI_TAB");
          return !(e && e[0] == '0');
        }();
        // two steps per A-fragment load (NS = 2): fewer shared-memory wavefronts but 168
        // registers -> 3 CTAs per SM; measured 676 ms against 646 ms for NS = 1 on C4, so
        // it stays an experiment (TQF_MVGBM_STEP_PAIRS=1)
        static const bool pair = [] {
          const char* e = std::getenv("TQF_MVGBM_STEP_PAIRS");
          return e && e[0] == '1';
        }();
        if (tab && a.ndtab != nullptr) {
          P.ndtab = a.ndtab;
          auto k_full = pair ? mvgbm_mma_kernel<true, true, 2> : mvgbm_mma_kernel<true, true, 1>;
          auto k_part = mvgbm_mma_kernel<false, true, 1>;
          TQF_CUDA_OK(cudaFuncSetAttribute(k_full, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kNdTabSmemBytes));
          TQF_CUDA_OK(cudaFuncSetAttribute(k_part, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kNdTabSmemBytes));
          if (a.dim == kMvDim)
            k_full<<<grid, kMmaWarps * 32, kNdTabSmemBytes, stream>>>(P);
          else
            k_part<<<grid, kMmaWarps * 32, kNdTabSmemBytes, stream>>>(P);
        } else if (a.dim == kMvDim) {
          mvgbm_mma_kernel<true, false, 1><<<grid, kMmaWarps * 32, 0, stream>>>(P);
        } else {
          mvgbm_mma_kernel<false, false, 1><<<grid, kMmaWarps * 32, 0, stream>>>(P);
        }
        TQF_CUDA_OK(cudaGetLastError());
        return TQF_OK;
      }
    }
    launch_split(P, grid, stream);
    TQF_CUDA_OK(cudaGetLastError());
    return TQF_OK;
  }
  int grid = static_cast<int>(P.num_chunks < static_cast<uint64_t>(a.max_grid)
                                  ? P.num_chunks
                                  : static_cast<uint64_t>(a.max_grid));
  if (grid < 1) grid = 1;
  *grid_out = grid;
  const size_t smem = static_cast<size_t>(DMAX) * kBlock * sizeof(Real);
  if constexpr (DMAX != kMvDim) {
    if (smem > 48 * 1024)
      TQF_CUDA_OK(cudaFuncSetAttribute(mvgbm_kernel<Real, DMAX>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    mvgbm_kernel<Real, DMAX><<<grid, kBlock, smem, stream>>>(P);
  }
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
