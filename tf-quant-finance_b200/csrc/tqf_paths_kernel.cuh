// The fused path kernel: one thread carries one path (or one antithetic pair)
// through every Euler step with its state in registers, drawing its normals
// in-kernel (Philox4x32-10 + Box-Muller, or Sobol XOR + inverse CDF) and
// either reducing payoffs (MODE_PRICE) or storing the state at the recorded
// steps (MODE_PATHS).
//
// Replaces the device work of models/euler_sampling.py:335-537 and
// models/utils.py:20-128 of the reference: there the [steps, N, dim] draws
// tensor is materialised and the [N, dim] state round-trips memory every step;
// here neither exists.
//
// Draw layout (models/utils.py:98-128): the flat element of (path p, step s,
// factor j) is p * (S_total * NF) + s * NF + j, for Philox the element index
// into tf.random.stateless_normal's stream, for Sobol dimension s * NF + j of
// point skip + 1 + p.
#pragma once

#include "tqf_common.cuh"

namespace tqf {

constexpr int kBlock = 128;       // threads per CTA == Sobol indices per chunk
constexpr int kLowBits = 7;       // log2(kBlock)
constexpr int kSobolTileDims = 512;  // Sobol dimensions staged in smem at once
constexpr int kWarps = kBlock / 32;

enum { MODE_PRICE = 0, MODE_PATHS = 1 };
enum { RNGK_PHILOX = 0, RNGK_SOBOL = 1, RNGK_DRAWS = 2 };

struct PayoffK {
  int32_t kind;
  int32_t component;
  int32_t transform;
  int32_t pad;
  double strike;
  double barrier;
  double scale;
};

template <typename Real>
struct KParams {
  // model
  const Real* coef;  // device [num_steps][NCOEF]
  int tables_in_smem;  // coef / record_slot staged in shared memory
  int num_steps;
  int num_steps_total;
  Real x0[2];
  // rng
  PhiloxKey key;
  PhiloxCtr ctr;
  uint64_t anti_half;        // N/2 for antithetic plans
  const uint32_t* sobol_v;   // device [S_total*NF][32], left aligned
  uint64_t first_index;      // Sobol: skip + 1 + path_offset ; else path_offset
  const Real* draws;         // device [N][S_total][NF]
  // work
  uint64_t path_offset;
  uint64_t path_count;
  uint64_t num_chunks;
  uint64_t chunk_base;       // first_index rounded down to a multiple of kBlock
  // MODE_PRICE
  int num_payoffs;
  int need_extrema;
  PayoffK pay[TQF_MAX_PAYOFFS];
  double* partials;          // device [gridDim.x][TQF_MAX_PAYOFFS][4]
  // MODE_PATHS
  const int* record_slot;    // device [num_steps + 1]
  Real* out;
  int64_t stride_path, stride_time, stride_dim;
};

// ------------------------------------------------------------- models -----
// coef columns: 0 = dt, 1 = sqrt(dt), then model specific.  Every step mirrors
// _euler_step (euler_sampling.py:513-537): dw = z sqrt_dt;
// x' = (x + dt a(t,x)) + S(t,x) dw with t = times[i+1].

template <typename R>
struct AffineModel1F {  // a = a0 + a1 x, S = b
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 5;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real* c) {
    const Real dw = z[0] * c[1];
    const Real dt_inc = c[0] * (c[2] + c[3] * x[0]);
    const Real dw_inc = c[4] * dw;
    x[0] = (x[0] + dt_inc) + dw_inc;
  }
};

template <typename R>
struct GbmModel1F {  // a = mu x, S = sigma x  (univariate_geometric_brownian_motion.py:127-153)
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 4;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real* c) {
    const Real dw = z[0] * c[1];
    const Real dt_inc = c[0] * (c[2] * x[0]);
    const Real dw_inc = (c[3] * x[0]) * dw;
    x[0] = (x[0] + dt_inc) + dw_inc;
  }
};

template <typename R>
struct LinearModel1F {  // x' = A x + B + C z  (HW exact OU step, vector_hull_white.py:738-767)
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 5;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real* c) {
    x[0] = (c[2] * x[0] + c[3]) + c[4] * z[0];
  }
};

template <typename R>
struct HestonEulerModel {  // heston/heston_model.py:143-173; state [X = log S, V]
  using Real = R;
  static constexpr int DIM = 2, NF = 2, NCOEF = 7;
  // c: dt, sqrt_dt, kappa, theta, volvol*rho, volvol*sqrt(1-rho^2), unused
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real* c) {
    const Real var = x[1];
    const Real vol = sqrt(fabs(var));
    const Real dw0 = z[0] * c[1];
    const Real dw1 = z[1] * c[1];
    const Real dx = c[0] * (Real(-0.5) * var);
    const Real dv = c[0] * (c[2] * (c[3] - var));
    x[0] = (x[0] + dx) + vol * dw0;
    x[1] = (var + dv) + ((c[4] * vol) * dw0 + (c[5] * vol) * dw1);
  }
};

// ------------------------------------------------------- normal streams ---
template <typename Real>
struct PhiloxStream;

template <>
struct PhiloxStream<double> {
  uint64_t group;
  double b0, b1;
  int pos;
  __device__ __forceinline__ void refill(const PhiloxKey& key, const PhiloxCtr& ctr) {
    const uint4 w = philox_group(ctr, key, group);
    box_muller(w.x, w.y, w.z, w.w, &b0, &b1);
    ++group;
  }
  __device__ __forceinline__ void init(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       uint64_t first_element) {
    group = first_element >> 1;
    refill(key, ctr);
    pos = static_cast<int>(first_element & 1);
  }
  __device__ __forceinline__ double next(const PhiloxKey& key, const PhiloxCtr& ctr) {
    if (pos == 2) {
      refill(key, ctr);
      pos = 0;
    }
    const double r = pos == 0 ? b0 : b1;
    ++pos;
    return r;
  }
};

template <>
struct PhiloxStream<float> {
  uint64_t group;
  float b0, b1, b2, b3;
  int pos;
  __device__ __forceinline__ void refill(const PhiloxKey& key, const PhiloxCtr& ctr) {
    const uint4 w = philox_group(ctr, key, group);
    box_muller(w.x, w.y, &b0, &b1);
    box_muller(w.z, w.w, &b2, &b3);
    ++group;
  }
  __device__ __forceinline__ void init(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       uint64_t first_element) {
    group = first_element >> 2;
    refill(key, ctr);
    pos = static_cast<int>(first_element & 3);
  }
  __device__ __forceinline__ float next(const PhiloxKey& key, const PhiloxCtr& ctr) {
    if (pos == 4) {
      refill(key, ctr);
      pos = 0;
    }
    const float r = pos == 0 ? b0 : (pos == 1 ? b1 : (pos == 2 ? b2 : b3));
    ++pos;
    return r;
  }
};

// ------------------------------------------------------------ payoffs -----
__device__ __forceinline__ double eval_payoff(const PayoffK& d, double x_final, double x_max,
                                              double x_min) {
  double f = x_final, fmax = x_max, fmin = x_min;
  if (d.transform == TQF_TRANSFORM_EXP) {
    f = exp(f);
    fmax = exp(fmax);
    fmin = exp(fmin);
  }
  double v;
  switch (d.kind) {
    case TQF_PAYOFF_CALL:
      v = f - d.strike > 0.0 ? f - d.strike : 0.0;
      break;
    case TQF_PAYOFF_PUT:
      v = d.strike - f > 0.0 ? d.strike - f : 0.0;
      break;
    case TQF_PAYOFF_UP_OUT_CALL:
      v = (f - d.strike > 0.0 && !(fmax > d.barrier)) ? f - d.strike : 0.0;
      break;
    case TQF_PAYOFF_UP_OUT_PUT:
      v = (d.strike - f > 0.0 && !(fmax > d.barrier)) ? d.strike - f : 0.0;
      break;
    case TQF_PAYOFF_DOWN_OUT_PUT:
      v = (d.strike - f > 0.0 && !(fmin < d.barrier)) ? d.strike - f : 0.0;
      break;
    case TQF_PAYOFF_DOWN_OUT_CALL:
      v = (f - d.strike > 0.0 && !(fmin < d.barrier)) ? f - d.strike : 0.0;
      break;
    default:  // TQF_PAYOFF_IDENTITY
      v = f;
      break;
  }
  return v * d.scale;
}

// -------------------------------------------------------------- kernel ----
template <class Model, int RNGK, bool ANTI, int MODE>
__global__ void __launch_bounds__(kBlock)
path_kernel(const KParams<typename Model::Real> P) {
  using Real = typename Model::Real;
  constexpr int DIM = Model::DIM, NF = Model::NF, NCOEF = Model::NCOEF;
  constexpr int NPATH = ANTI ? 2 : 1;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: coef [S][NCOEF] Real | record_slot [S+1] int | sobol high [T] u32
  //         | sobol low [T][8] u32 | accumulators [kWarps][8][3] double
  Real* s_coef = reinterpret_cast<Real*>(smem_raw);
  size_t off = 0;
  if (P.tables_in_smem) {
    off = static_cast<size_t>(P.num_steps) * NCOEF * sizeof(Real);
    off = (off + 15) & ~static_cast<size_t>(15);
  }
  int* s_rec = reinterpret_cast<int*>(smem_raw + off);
  if (MODE == MODE_PATHS && P.tables_in_smem)
    off += ((static_cast<size_t>(P.num_steps) + 1) * sizeof(int) + 15) & ~static_cast<size_t>(15);
  uint32_t* s_high = reinterpret_cast<uint32_t*>(smem_raw + off);
  if (RNGK == RNGK_SOBOL) off += kSobolTileDims * sizeof(uint32_t);
  uint4* s_low = reinterpret_cast<uint4*>(smem_raw + off);
  if (RNGK == RNGK_SOBOL) off += static_cast<size_t>(kSobolTileDims) * 8 * sizeof(uint32_t);
  double* s_acc = reinterpret_cast<double*>(smem_raw + off);

  const int tid = threadIdx.x;
  const Real* coef_tab = P.coef;
  const int* rec_tab = P.record_slot;
  if (P.tables_in_smem) {
    for (int i = tid; i < P.num_steps * NCOEF; i += kBlock) s_coef[i] = P.coef[i];
    coef_tab = s_coef;
    if (MODE == MODE_PATHS) {
      for (int i = tid; i <= P.num_steps; i += kBlock) s_rec[i] = P.record_slot[i];
      rec_tab = s_rec;
    }
  }
  if (MODE == MODE_PRICE) {
    for (int i = tid; i < kWarps * TQF_MAX_PAYOFFS * 3; i += kBlock) s_acc[i] = 0.0;
  }
  __syncthreads();

  // Sobol: masks of this thread's low index bits (constant over chunks).
  uint32_t lowmask[kLowBits];
#pragma unroll
  for (int b = 0; b < kLowBits; ++b) lowmask[b] = 0u - ((static_cast<uint32_t>(tid) >> b) & 1u);

  constexpr int TILE_STEPS = (kSobolTileDims / NF) > 0 ? (kSobolTileDims / NF) : 1;
  const uint64_t stream_stride = static_cast<uint64_t>(P.num_steps_total) * NF;

  for (uint64_t chunk = blockIdx.x; chunk < P.num_chunks; chunk += gridDim.x) {
    const uint64_t index = P.chunk_base + chunk * kBlock + tid;  // global unit / Sobol index
    const bool valid = index >= P.first_index && index < P.first_index + P.path_count;
    const uint64_t local = index - P.first_index;              // row inside the shard
    const uint64_t unit = P.path_offset + local;               // global path (or pair) number

    Real x[NPATH][DIM];
#pragma unroll
    for (int a = 0; a < NPATH; ++a)
#pragma unroll
      for (int j = 0; j < DIM; ++j) x[a][j] = P.x0[j];
    Real xmax[NPATH][DIM], xmin[NPATH][DIM];
#pragma unroll
    for (int a = 0; a < NPATH; ++a)
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        xmax[a][j] = x[a][j];
        xmin[a][j] = x[a][j];
      }

    PhiloxStream<Real> stream;
    if (RNGK == RNGK_PHILOX) stream.init(P.key, P.ctr, valid ? unit * stream_stride : 0);
    const Real* my_draws = nullptr;
    if (RNGK == RNGK_DRAWS) my_draws = P.draws + (valid ? unit : 0) * stream_stride;

    if (MODE == MODE_PATHS) {
      const int slot = rec_tab[0];
      if (slot >= 0 && valid) {
#pragma unroll
        for (int a = 0; a < NPATH; ++a)
#pragma unroll
          for (int j = 0; j < DIM; ++j)
            P.out[static_cast<int64_t>(local + a * P.anti_half) * P.stride_path +
                  slot * P.stride_time + j * P.stride_dim] = x[a][j];
      }
    }

    for (int s0 = 0; s0 < P.num_steps; s0 += TILE_STEPS) {
      const int s1 = min(P.num_steps, s0 + TILE_STEPS);
      if (RNGK == RNGK_SOBOL) {
        // Stage the direction numbers of dimensions [s0*NF, s1*NF): the XOR of
        // the chunk's common high index bits, and the kLowBits low columns.
        __syncthreads();
        const uint32_t high_bits = static_cast<uint32_t>((P.chunk_base + chunk * kBlock) >> kLowBits);
        for (int dd = tid; dd < (s1 - s0) * NF; dd += kBlock) {
          const uint32_t* v = P.sobol_v + (static_cast<size_t>(s0) * NF + dd) * 32;
          const uint4 l0 = *reinterpret_cast<const uint4*>(v);
          const uint4 l1 = *reinterpret_cast<const uint4*>(v + 4);
          s_low[2 * dd] = l0;
          s_low[2 * dd + 1] = l1;
          uint32_t h = 0;
          uint32_t hb = high_bits;
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
        Real z[NF];
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          if (RNGK == RNGK_PHILOX) {
            z[j] = stream.next(P.key, P.ctr);
          } else if (RNGK == RNGK_SOBOL) {
            const int dd = (s - s0) * NF + j;
            const uint4 l0 = s_low[2 * dd];
            const uint4 l1 = s_low[2 * dd + 1];
            uint32_t xb = s_high[dd];
            xb ^= l0.x & lowmask[0];
            xb ^= l0.y & lowmask[1];
            xb ^= l0.z & lowmask[2];
            xb ^= l0.w & lowmask[3];
            xb ^= l1.x & lowmask[4];
            xb ^= l1.y & lowmask[5];
            xb ^= l1.z & lowmask[6];
            z[j] = ndtri(RealTraits<Real>::sobol_uniform(xb));
          } else {
            z[j] = my_draws[static_cast<size_t>(s) * NF + j];
          }
        }
        const Real* c = coef_tab + s * NCOEF;
        Model::step(x[0], z, c);
        if (ANTI) {
          Real zm[NF];
#pragma unroll
          for (int j = 0; j < NF; ++j) zm[j] = -z[j];
          Model::step(x[NPATH - 1], zm, c);
        }
        if (MODE == MODE_PRICE) {
          if (P.need_extrema) {
#pragma unroll
            for (int a = 0; a < NPATH; ++a)
#pragma unroll
              for (int j = 0; j < DIM; ++j) {
                xmax[a][j] = x[a][j] > xmax[a][j] ? x[a][j] : xmax[a][j];
                xmin[a][j] = x[a][j] < xmin[a][j] ? x[a][j] : xmin[a][j];
              }
          }
        } else {
          const int slot = rec_tab[s + 1];
          if (slot >= 0 && valid) {
#pragma unroll
            for (int a = 0; a < NPATH; ++a)
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                P.out[static_cast<int64_t>(local + a * P.anti_half) * P.stride_path +
                      slot * P.stride_time + j * P.stride_dim] = x[a][j];
          }
        }
      }
    }

    if (MODE == MODE_PRICE) {
      const int warp = tid >> 5, lane = tid & 31;
      for (int q = 0; q < P.num_payoffs; ++q) {
        const PayoffK& d = P.pay[q];
        double sum = 0.0, sq = 0.0, bad = 0.0;
        if (valid) {
#pragma unroll
          for (int a = 0; a < NPATH; ++a) {
            double xf = 0.0, xa = 0.0, xi = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; ++j) {
              if (j == d.component) {
                xf = static_cast<double>(x[a][j]);
                xa = static_cast<double>(xmax[a][j]);
                xi = static_cast<double>(xmin[a][j]);
              }
            }
            const double v = eval_payoff(d, xf, xa, xi);
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
          double* acc = s_acc + (warp * TQF_MAX_PAYOFFS + q) * 3;
          acc[0] += sum;
          acc[1] += sq;
          acc[2] += bad;
        }
      }
    }
  }

  if (MODE == MODE_PRICE) {
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

// Deterministic final reduction of the per-CTA partials.
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int num_blocks,
                                       int num_payoffs, double* __restrict__ sums);

template <typename Real>
size_t path_kernel_smem(int ncoef, int num_steps, int rngk, int mode, bool tables_in_smem) {
  size_t off = 0;
  if (tables_in_smem) {
    off = static_cast<size_t>(num_steps) * ncoef * sizeof(Real);
    off = (off + 15) & ~static_cast<size_t>(15);
    if (mode == MODE_PATHS) off += ((static_cast<size_t>(num_steps) + 1) * sizeof(int) + 15) & ~static_cast<size_t>(15);
  }
  if (rngk == RNGK_SOBOL) off += kSobolTileDims * sizeof(uint32_t) + static_cast<size_t>(kSobolTileDims) * 8 * sizeof(uint32_t);
  off += static_cast<size_t>(kWarps) * TQF_MAX_PAYOFFS * 3 * sizeof(double);
  return off;
}

// Launches the right instantiation for (rng kind, antithetic, mode).
template <class Model>
int launch_path_kernel(int rngk, bool anti, int mode, int grid, size_t smem,
                       const KParams<typename Model::Real>& P, cudaStream_t stream) {
#define TQF_LAUNCH(RK, AN, MD)                                                         \
  do {                                                                                 \
    auto kern = path_kernel<Model, RK, AN, MD>;                                        \
    if (smem > 48 * 1024)                                                              \
      TQF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       static_cast<int>(smem)));                       \
    kern<<<grid, kBlock, smem, stream>>>(P);                                           \
    TQF_CUDA_OK(cudaGetLastError());                                                   \
    return TQF_OK;                                                                     \
  } while (0)
  if (mode == MODE_PRICE) {
    if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PRICE);
    if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PRICE);
    if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PRICE);
    if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PRICE);
  } else {
    if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PATHS);
    if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PATHS);
    if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PATHS);
    if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PATHS);
  }
#undef TQF_LAUNCH
  set_error("unsupported rng / mode combination");
  return TQF_ERR_UNSUPPORTED;
}

}  // namespace tqf
