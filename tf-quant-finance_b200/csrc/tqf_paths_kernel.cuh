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

#include <type_traits>

#include <cstdlib>

#include "tqf_common.cuh"

namespace tqf {

#ifndef TQF_MIN_BLOCKS
#define TQF_MIN_BLOCKS 3
#endif
constexpr int kBlock = 128;       // threads per CTA == Sobol indices per chunk
constexpr int kLowBits = 7;       // log2(kBlock)
#ifndef TQF_SOBOL_TILE
#define TQF_SOBOL_TILE 256
#endif
constexpr int kSobolTileDims = TQF_SOBOL_TILE;  // Sobol dimensions staged in smem at once
constexpr int kWarps = kBlock / 32;
constexpr int kMaxPPT = 8;      // most paths carried by one thread

// MODE_PRICE_EXTREMA: MODE_PRICE that also tracks the running extrema of the monitored
// component (barrier payoffs); a mode of its own so that the plain pricing kernels carry
// neither the registers nor the per-step test.
// MODE_PRICE_BRIDGE: MODE_PRICE_EXTREMA plus the Brownian-bridge no-touch products.
enum { MODE_PRICE = 0, MODE_PATHS = 1, MODE_PRICE_EXTREMA = 2, MODE_PRICE_BRIDGE = 3 };
enum { RNGK_PHILOX = 0, RNGK_SOBOL = 1, RNGK_DRAWS = 2 };

struct PayoffK {
  int32_t kind;
  int32_t component;
  int32_t transform;
  int32_t step;   // evaluated on the state after this many steps
  double strike;
  double barrier;
  double scale;
  int32_t tangent;   // TQF_PAYOFF_*_TANGENT: state component of the tangent
  int32_t bridge;    // barrier payoffs: multiply by the bridge no-touch probability
};

// Device-side table of one Hull-White swaption payoff (TQF_PAYOFF_HW_SWAPTION):
//   P(t_e, T_j) = exp(k_j - G_j x),  payoff = scale max(+-DF (1 - sum_j coef_j P_j), 0)
struct SwaptionK {
  int32_t num_payments;
  int32_t is_payer;
  int32_t num_factors;   // g[j * num_factors + i] multiplies state component i
  int32_t pad;
  double g[TQF_MAX_SWAPTION_PAYMENTS];
  double k[TQF_MAX_SWAPTION_PAYMENTS];
  double coef[TQF_MAX_SWAPTION_PAYMENTS];
};

template <typename Real>
struct KParams {
  // model
  const Real* coef;  // device [num_steps][NCOEF]
  int tables_in_smem;  // coef / record_slot staged in shared memory
  int num_steps;
  int num_steps_total;
  Real x0[4];
  const Real* x0_paths;      // device [N][DIM] per-path initial states, or null
  uint64_t x0_half;          // antithetic plans: row of the partner = row + x0_half (N / 2)
  // rng
  PhiloxKey key;
  PhiloxCtr ctr;
  uint64_t anti_half;        // N/2 for antithetic plans
  const uint32_t* sobol_v;   // device [S_total*NF][32], left aligned
  const double* logtab;      // device-global log table (tqf_math.cuh), double Sobol only
  int sobol_hi[8];           // 8 x 0x41400000 (see sobol_normals): separate params so that
                             // each lives in its own register
  int sobol_clamp;           // float32 Sobol: u == 1.0 -> largest float below 1 (non-reference mode)
  uint64_t first_index;      // Sobol: skip + 1 + path_offset ; else path_offset
  const Real* draws;         // device [N][S_total][NF]
  // work
  uint64_t path_offset;
  uint64_t unit_stride, unit_offset;  // draw unit of path p: p * unit_stride + unit_offset
  uint64_t path_count;
  uint64_t num_chunks;
  uint64_t chunk_base;       // first_index rounded down to a multiple of kBlock
  // MODE_PRICE
  int num_payoffs;
  int need_extrema;          // bit 0: running max, bit 1: running min of `monitor`
  int monitor;               // state component the barrier payoffs watch
  // Brownian-bridge correction of the barrier payoffs (continuous monitoring,
  // black_scholes/brownian_bridge.py:118-196): bit 0 = an upper, bit 1 = a lower
  // barrier at bridge_up / bridge_dn (state space, component 0)
  int bridge;
  double bridge_up, bridge_dn;
  PayoffK pay[TQF_MAX_PAYOFFS];
  double* partials;          // device [gridDim.x][TQF_MAX_PAYOFFS][4]
  const SwaptionK* swaptions; // device [num_payoffs] (HW swaption payoffs only)
  // MODE_PATHS
  const int* record_slot;    // device [num_steps + 1]
  Real* out;
  int64_t stride_path, stride_time, stride_dim;
  int store_exp;             // store exp(state) (log-space models feeding LSM)
  // optional column sums of the stored values (the basis-centring means of the
  // Longstaff-Schwartz passes, lsm.py:110-111, for free while the paths are written)
  double* colsum_partials;   // device [gridDim.x][colsum_cols] or null
  int colsum_cols;           // number of time slots * DIM
};

// v[comp] for a register-resident state vector.  Written with opaque `selp`s:
// a plain `j == comp ? v[j] : r` chain is turned into an indexed load by the
// compiler, which moves the whole state array to local memory (an LDL/STL pair
// per path and step in the hot loop).
__device__ __forceinline__ double selp_real(double a, double b, int take_a) {
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}"
      : "=d"(r) : "d"(a), "d"(b), "r"(take_a));
  return r;
}
__device__ __forceinline__ float selp_real(float a, float b, int take_a) {
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}"
      : "=f"(r) : "f"(a), "f"(b), "r"(take_a));
  return r;
}
// ------------------------------------------------------------- models -----
// coef columns: 0 = dt, 1 = sqrt(dt), then model specific.  Every step mirrors
// _euler_step (euler_sampling.py:513-537): dw = z sqrt_dt;
// x' = (x + dt a(t,x)) + S(t,x) dw with t = times[i+1].

template <typename R>
struct AffineModel1F {  // a = a0 + a1 x, S = b0 + b1 x
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 6;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    const Real dw = z[0] * c[1];
    const Real dt_inc = c[0] * (c[2] + c[3] * x[0]);
    const Real dw_inc = (c[4] + c[5] * x[0]) * dw;
    x[0] = (x[0] + dt_inc) + dw_inc;
  }
  __device__ static __forceinline__ Real bridge_var(const Real (&x)[DIM], const Real (&c)[NCOEF]) {
    const Real vol = c[4] + c[5] * x[0];
    return vol * vol * c[0];
  }
};

// Generic affine Ito process of dimension D (2..4):
//   a(t, x) = a0(t) + A1(t) x,   S(t, x) = B(t)   (state-independent volatility)
// coef: dt, sqrt_dt, a0[D], A1[D][D] row-major, B[D][D] row-major.
template <typename R, int D>
struct AffineModelND {
  using Real = R;
  static constexpr int DIM = D, NF = D, NCOEF = 2 + D + 2 * D * D;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    Real dw[D], xn[D];
#pragma unroll
    for (int j = 0; j < D; ++j) dw[j] = z[j] * c[1];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      Real drift = c[2 + i];
      Real diff = 0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        drift = fma(c[2 + D + i * D + j], x[j], drift);
        diff = fma(c[2 + D + D * D + i * D + j], dw[j], diff);
      }
      xn[i] = (x[i] + c[0] * drift) + diff;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
  }
};
template <typename R> using AffineModel2D = AffineModelND<R, 2>;
template <typename R> using AffineModel3D = AffineModelND<R, 3>;
template <typename R> using AffineModel4D = AffineModelND<R, 4>;

// AffineModel1F with its pathwise tangents (forward-mode sensitivities, the
// Jacobian-carrying loop of euler_sampling.py:393-402 / custom_loops.py:20-215):
//   y = dx/dx0:      y' = y + y (dt a1 + b1 dw)
//   v = dx/dtheta:   v' = v + v (dt a1 + b1 dw) + dt (da0 + da1 x) + (db0 + db1 x) dw
// with the PRE-step x on the right-hand sides (differentiating _euler_step).
template <typename R>
struct TangentAffine1FModel {
  using Real = R;
  static constexpr int DIM = 3, NF = 1, NCOEF = 10;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    const Real dw = z[0] * c[1];
    const Real xs = x[0];
    const Real g = fma(c[5], dw, c[0] * c[3]);           // d(increment)/dx
    const Real dt_inc = c[0] * (c[2] + c[3] * xs);
    const Real dw_inc = (c[4] + c[5] * xs) * dw;
    x[0] = (xs + dt_inc) + dw_inc;
    x[2] = fma(x[2], g, x[2]) + (c[0] * (c[6] + c[7] * xs) + (c[8] + c[9] * xs) * dw);
    x[1] = fma(x[1], g, x[1]);
  }
};

// Milstein step in one dimension (_milstein_1d, milstein_sampling.py:565-575):
//   x' = x + dt a + b dw + (b b') (dw^2 - dt) / 2,  a = a0 + a1 x, b = b0 + b1 x, b' = b1,
// with the reference's grouping ((x + dt_inc) + dw_inc) + hot_inc.
template <typename R>
struct MilsteinAffine1FModel {
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 6;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    const Real dw = z[0] * c[1];
    const Real vol = c[4] + c[5] * x[0];
    const Real dt_inc = c[0] * (c[2] + c[3] * x[0]);
    const Real dw_inc = vol * dw;
    const Real hot_inc = ((vol * c[5]) * (dw * dw - c[0])) / Real(2);
    x[0] = ((x[0] + dt_inc) + dw_inc) + hot_inc;
  }
};

template <typename R>
struct GbmModel1F {  // a = mu x, S = sigma x  (univariate_geometric_brownian_motion.py:127-153)
  using Real = R;
  static constexpr int DIM = 1, NF = 1, NCOEF = 4;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
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
                                              const Real (&c)[NCOEF]) {
    x[0] = (c[2] * x[0] + c[3]) + c[4] * z[0];
  }
  __device__ static __forceinline__ Real bridge_var(const Real (&)[DIM], const Real (&c)[NCOEF]) {
    return c[4] * c[4];
  }
};

template <typename R>
struct HullWhite1FModel {
  // Exact OU step of the one-factor Hull-White model
  // (vector_hull_white.py:738-767) with the path integral of the short rate
  // carried as a second state component (hjm/swaption_util.py:126-136):
  //   x' = A x + B + C z ;  I' = I + W (x' + f(0, t')) = I + W x' + WF
  using Real = R;
  static constexpr int DIM = 2, NF = 1, NCOEF = 5;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    x[0] = fma(c[2], z[0], fma(c[0], x[0], c[1]));
    x[1] = fma(c[3], x[0], x[1] + c[4]);
  }
};

// Gaussian / quasi-Gaussian HJM with deterministic volatility (TQF_MODEL_HJM, see
// tqf.h): F Markov factors and the short-rate integral; NFS normals are consumed
// per step (the reference's Wiener process also drives the zero-volatility vec(y)
// components of the quasi-Gaussian state), the first F act.
template <typename R, int F, int NFS>
struct HjmModel {
  using Real = R;
  static constexpr int DIM = F + 1, NF = NFS, NCOEF = 5 + 2 * F + F * F;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    Real dw[F], xn[F];
    Real pre = 0, post = 0;
#pragma unroll
    for (int j = 0; j < F; ++j) dw[j] = z[j] * c[1];
#pragma unroll
    for (int i = 0; i < F; ++i) {
      const Real drift = c[2 + i] - c[2 + F + i] * x[i];
      Real diff = 0;
#pragma unroll
      for (int j = 0; j < F; ++j) diff = fma(c[2 + 2 * F + i * F + j], dw[j], diff);
      xn[i] = (x[i] + c[0] * drift) + diff;
      pre += x[i];
      post += xn[i];
    }
#pragma unroll
    for (int i = 0; i < F; ++i) x[i] = xn[i];
    x[F] = x[F] + (c[2 + 2 * F + F * F] * pre + c[3 + 2 * F + F * F] * post + c[4 + 2 * F + F * F]);
  }
};
template <typename R> using HjmModel11 = HjmModel<R, 1, 1>;
template <typename R> using HjmModel12 = HjmModel<R, 1, 2>;
template <typename R> using HjmModel22 = HjmModel<R, 2, 2>;
template <typename R> using HjmModel26 = HjmModel<R, 2, 6>;
template <typename R> using HjmModel33 = HjmModel<R, 3, 3>;

template <typename R>
struct HestonEulerModel {  // heston/heston_model.py:143-173; state [X = log S, V]
  using Real = R;
  static constexpr int DIM = 2, NF = 2, NCOEF = 6;
  // c: sqrt_dt, -dt/2, dt*kappa, theta, volvol*rho*sqrt_dt, volvol*sqrt(1-rho^2)*sqrt_dt
  // (per-step products formed once on the host).  Same update as _euler_step
  // with the Heston closures,
  //   X' = X + dt (-V/2) + sqrt|V| sqrt_dt z0
  //   V' = V + dt kappa (theta - V) + volvol sqrt|V| sqrt_dt (rho z0 + sqrt(1-rho^2) z1),
  // regrouped into 8 FMA-pipe operations + the square root.
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    const Real var = x[1];
    const Real vol = sqrt_abs(var);
    x[0] = fma(vol, z[0] * c[0], fma(c[1], var, x[0]));
    x[1] = fma(vol, fma(c[5], z[1], c[4] * z[0]), fma(c[2], c[3] - var, var));
  }
  // log-spot increment: sqrt|V| sqrt_dt z0
  __device__ static __forceinline__ Real bridge_var(const Real (&x)[DIM], const Real (&c)[NCOEF]) {
    return (x[1] < Real(0) ? -x[1] : x[1]) * (c[0] * c[0]);
  }
  __device__ static __forceinline__ double sqrt_abs(double v) {
    // |V| == 0 would make rsqrt infinite: nudge it (the addend is absorbed otherwise).
    return fm::sqrt_pos(fabs(v) + 1e-300);  // absorbed unless V == 0
  }
  __device__ static __forceinline__ float sqrt_abs(float v) { return sqrtf(fabsf(v)); }
};

// HestonEulerModel with the pathwise tangents of (X, V) with respect to ONE scalar
// parameter p (forward-mode sensitivities: what the reference obtains by differentiating
// the Euler loop with `watch_params`, euler_sampling.py:393-402).  State
// [X, V, dX/dp, dV/dp]; with s = sqrt|V|, ds = sign(V) (dV/dp) / (2 s) and the PRE-step
// state on the right-hand sides (differentiating _euler_step with the closures of
// heston/heston_model.py:143-173):
//   dX/dp' = dX/dp - dt (dV/dp) / 2 + ds dw0
//   dV/dp' = dV/dp + dt (dkappa (theta - V) + kappa (dtheta - dV/dp))
//            + (dxi s + xi ds) (rho dw0 + rhobar dw1) + xi s (drho dw0 + drhobar dw1).
// c: dt, sqrt_dt, kappa, theta, xi, rho, rhobar, dkappa, dtheta, dxi, drho, drhobar.
// The primal follows the reference grouping (x + dt drift) + vol . dw.
template <typename R>
struct TangentHestonModel {
  using Real = R;
  static constexpr int DIM = 4, NF = 2, NCOEF = 12;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    const Real dw0 = z[0] * c[1], dw1 = z[1] * c[1];
    const Real v = x[1], vt = x[3];
    const Real s = sqrt(v < Real(0) ? -v : v);
    const Real ds = s > Real(0) ? (v < Real(0) ? -vt : vt) / (Real(2) * s) : Real(0);
    const Real w = c[5] * dw0 + c[6] * dw1;
    const Real wp = c[9 + 1] * dw0 + c[9 + 2] * dw1;
    const Real xt = (x[2] + c[0] * (Real(-0.5) * vt)) + ds * dw0;
    const Real vtn = (vt + c[0] * (c[7] * (c[3] - v) + c[2] * (c[8] - vt))) +
                     ((c[9] * s + c[4] * ds) * w + (c[4] * s) * wp);
    x[0] = (x[0] + c[0] * (Real(-0.5) * v)) + s * dw0;
    x[1] = (v + c[0] * (c[2] * (c[3] - v))) + (c[4] * s) * w;
    x[2] = xt;
    x[3] = vtn;
  }
};

template <typename R>
struct HestonQeModel {
  // Andersen's Quadratic-Exponential step (heston/heston_model.py:402-431,
  // 522-572): z[0] drives the variance, z[1] the log-spot.  Per-step constants
  // from the host: active (dt > tolerance), e = exp(-kappa dt), theta,
  // s^2 = c_s1 V + c_s0, and k0..k4 of the log-spot update.
  using Real = R;
  static constexpr int DIM = 2, NF = 2, NCOEF = 10;
  __device__ static __forceinline__ void step(Real (&x)[DIM], const Real (&z)[NF],
                                              const Real (&c)[NCOEF]) {
    if (c[0] == Real(0)) return;  // zero-length step: consumes its draws only
    step_impl(x, z, c);
  }
  // Reference arithmetic op for op (IEEE divisions and square roots of libdevice).
  template <typename T>
  __device__ static __forceinline__ void step_reference(T (&x)[DIM], const T (&z)[NF],
                                                        const T (&c)[NCOEF]) {
    const T v = x[1];
    const T m = c[2] + (v - c[2]) * c[1];
    const T s2 = fma(v, c[3], c[4]);
    const T psi = s2 / (m * m);
    T vn;
    if (psi < T(1.5)) {
      const T psi_inv = T(2) / psi;
      const T b2 = psi_inv - T(1) + sqrt(psi_inv * (psi_inv - T(1)));
      const T a = m / (T(1) + b2);
      const T t = sqrt(b2) + z[0];
      vn = a * (t * t);
    } else {
      const T p = (psi - T(1)) / (psi + T(1));
      const T beta = (T(1) - p) / m;
      const T u = T(0.5) * (T(1) + erf(z[0] * T(0.70710678118654752440)));
      vn = u > p ? (log(T(1) - p) - log(T(1) - u)) / beta : T(0);
    }
    x[0] = (((x[0] + c[5]) + c[6] * v) + c[7] * vn) + sqrt(c[8] * v + c[9] * vn) * z[1];
    x[1] = vn;
  }
  __device__ static __forceinline__ void step_impl(float (&x)[DIM], const float (&z)[NF],
                                                   const float (&c)[NCOEF]) {
    step_reference<float>(x, z, c);
  }
  // float64 (see step_batch): the quadratic branch for one path.
  __device__ static __forceinline__ void step_impl(double (&x)[DIM], const double (&z)[NF],
                                                   const double (&c)[NCOEF]) {
    double xs[1][1][DIM] = {{{x[0], x[1]}}};
    const double zs[1][NF] = {{z[0], z[1]}};
    step_batch<1, 1, 0>(xs, zs, c, 1.0);
    x[0] = xs[0][0][0];
    x[1] = xs[0][0][1];
  }

  // The QE step of the N paths a thread carries (float64), half H of each
  // antithetic pair, normals sign * z:
  //  * the quadratic branch (psi < 1.5, ~97% of the path-steps of config C2) runs for
  //    all N side by side on the hand-written reciprocal / square root (5-6 FP64
  //    instructions each, <= 1 ulp, no special-case code): two reciprocals and three
  //    square roots per step instead of three IEEE divisions and square roots;
  //  * the paths that need the exponential branch (psi >= 1.5: variance near zero) or
  //    have degenerate inputs are then served ONE PER ITERATION of a compacting loop
  //    (a lane with one such path runs the ~300-instruction reference arithmetic once;
  //    a branch per path would run it once per path for the whole warp).
  static constexpr bool kBatchStep = sizeof(R) == 8;
  template <int N, int NP, int H>
  __device__ static __forceinline__ void step_batch(double (&x)[N][NP][DIM],
                                                    const double (&z)[N][NF],
                                                    const double (&c)[NCOEF], double sign) {
    if (c[0] == 0.0) return;
    double v[N], m[N], s2[N], vn[N];
    unsigned slow = 0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      v[a] = x[a][H][1];
      m[a] = c[2] + (v[a] - c[2]) * c[1];
      s2[a] = fma(v[a], c[3], c[4]);
      const double m2 = m[a] * m[a];
      // psi = s2 / m2 < 1.5  <=>  s2 < 1.5 m2 for positive m2
      const bool quad = m[a] > 1e-150 && s2[a] > 1e-300 && s2[a] < 1.5 * m2 && m2 < 1e150;
      slow |= quad ? 0u : (1u << a);
      const double q = (m2 + m2) * fm::rcp_pos(quad ? s2[a] : 1.0);     // 2 / psi  (> 4/3)
      const double t1 = q - 1.0;
      const double b2 = t1 + fm::sqrt_pos(q * t1);
      const double al = m[a] * fm::rcp_pos(1.0 + b2);
      const double t = fm::sqrt_pos(b2) + sign * z[a][0];
      vn[a] = al * (t * t);
    }
    while (slow) {
      const int a = __ffs(slow) - 1;
      slow &= slow - 1;
      double ms = m[0], ss = s2[0], zs = z[0][0];
#pragma unroll
      for (int k = 1; k < N; ++k) {
        ms = selp_real(m[k], ms, k == a ? 1 : 0);
        ss = selp_real(s2[k], ss, k == a ? 1 : 0);
        zs = selp_real(z[k][0], zs, k == a ? 1 : 0);
      }
      // reference arithmetic (heston_model.py:522-551)
      const double psi = ss / (ms * ms);
      double r;
      if (psi < 1.5) {
        const double psi_inv = 2.0 / psi;
        const double b2 = psi_inv - 1.0 + sqrt(psi_inv * (psi_inv - 1.0));
        const double t = sqrt(b2) + sign * zs;
        r = ms / (1.0 + b2) * (t * t);
      } else {
        const double p = (psi - 1.0) / (psi + 1.0);
        const double beta = (1.0 - p) / ms;
        const double u = 0.5 * (1.0 + erf(sign * zs * 0.70710678118654752440));
        r = u > p ? (log(1.0 - p) - log(1.0 - u)) / beta : 0.0;
      }
#pragma unroll
      for (int k = 0; k < N; ++k) vn[k] = selp_real(r, vn[k], k == a ? 1 : 0);
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
      const double arg = fma(c[8], v[a], c[9] * vn[a]);
      // (the addend keeps rsqrt finite when v = vn = 0 and is absorbed otherwise)
      const double sq = arg > 0.0 ? fm::sqrt_pos(arg + 1e-300) : sqrt(arg);
      x[a][H][0] = (((x[a][H][0] + c[5]) + c[6] * v[a]) + c[7] * vn[a]) + sq * (sign * z[a][1]);
      x[a][H][1] = vn[a];
    }
  }
};

// Variance of the increment of state component 0 over one step, given the state
// BEFORE the step: the `variance` argument of brownian_bridge_single
// (black_scholes/brownian_bridge.py:118-196).  Models that define bridge_var can
// price continuously monitored barriers (TQF payoff flag `brownian_bridge`).
template <class M, class = void>
struct HasBridgeVar {
  static constexpr bool value = false;
};
template <class M>
struct HasBridgeVar<M, decltype(void(&M::bridge_var))> {
  static constexpr bool value = true;
};

// Models that step all the paths of a thread at once define kBatchStep = true.
template <class M, class = void>
struct HasBatchStep {
  static constexpr bool value = false;
};
template <class M>
struct HasBatchStep<M, decltype(void(M::kBatchStep))> {
  static constexpr bool value = M::kBatchStep;
};

// ------------------------------------------------------- normal streams ---
// PPT independent Philox streams held by one thread (one per path it carries);
// all of them sit at the same position inside their group, so the Box-Muller
// of the PPT groups is evaluated side by side (tqf_math.cuh).
template <typename Real, int PPT>
struct PhiloxStreamV;

// BoxMullerDouble (random_distributions.h) of PPT Philox groups side by side:
//   u1 = max(U(x0, x1), 1e-7), v = 2 pi U(x2, x3), r = sqrt(-2 ln u1) -> (r sin v, r cos v).
// -ln u1 comes from the table logarithm (tqf_logtab0.inc, MID-free: 8 FP64
// instructions instead of 18 + MUFU + I2F); its absolute error of a few 1e-19 near
// u1 = 1 and 1e-15 near 2^-14 keeps r within 1e-14 relative wherever r > 1e-5.
// Arguments below 2^-14 (6e-5 of the draws, u1 = 0 and the 1e-7 clamp included)
// take the full-precision logarithm in a rarely entered branch.
template <int PPT, class Tab>
__device__ __forceinline__ void philox_box_muller_v(const PhiloxKey& key, const PhiloxCtr& ctr,
                                                    const Tab& tab, uint64_t (&group)[PPT],
                                                    double (&z0)[PPT], double (&z1)[PPT]) {
  double u1[PPT], v1[PPT], w[PPT], sn[PPT], cs[PPT];
#pragma unroll
  for (int a = 0; a < PPT; ++a) {
    const uint4 q = philox_group_nowrap(ctr, key, group[a]);   // group[] = counter low half
    ++group[a];
    u1[a] = uint64_to_double(q.x, q.y);
    v1[a] = 6.283185307179586476925286766559 * uint64_to_double(q.z, q.w);
  }
  uint32_t hmin;
  fm::neg_log_mid_tab_v<PPT>(tab, u1, w, &hmin);
  if (hmin < 0x3f100000u) {                       // some u1 < 2^-14: outside the table
#pragma unroll
    for (int a = 0; a < PPT; ++a)
      if (static_cast<uint32_t>(__double2hiint(u1[a])) < 0x3f100000u) {
        const double u = u1[a] < 1.0e-7 ? 1.0e-7 : u1[a];
        w[a] = -fm::log_pos(u);
      }
  }
  fm::sincos_2pi_v<PPT>(tab, v1, sn, cs);
#pragma unroll
  for (int a = 0; a < PPT; ++a) {
    const double r = fm::sqrt_pos(w[a] + w[a]);
    z0[a] = sn[a] * r;
    z1[a] = cs[a] * r;
  }
}

template <int PPT>
struct PhiloxStreamV<double, PPT> {
  uint64_t group[PPT];
  double b0[PPT], b1[PPT];
  int pos;
  template <class Tab>
  __device__ __forceinline__ void refill(const PhiloxKey& key, const PhiloxCtr& ctr,
                                         const Tab& tab) {
    philox_box_muller_v<PPT>(key, ctr, tab, group, b0, b1);
  }
  template <class Tab>
  __device__ __forceinline__ void init(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       const Tab& tab,
                                       const uint64_t (&first_element)[PPT]) {
#pragma unroll
    for (int a = 0; a < PPT; ++a) group[a] = philox_base_lo(ctr) + (first_element[a] >> 1);
    refill(key, ctr, tab);
    pos = static_cast<int>(first_element[0] & 1);
  }
  template <class Tab>
  __device__ __forceinline__ void next(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       const Tab& tab, double (&z)[PPT]) {
    if (pos == 2) {
      refill(key, ctr, tab);
      pos = 0;
    }
#pragma unroll
    for (int a = 0; a < PPT; ++a) z[a] = pos == 0 ? b0[a] : b1[a];
    ++pos;
  }
};

template <int PPT>
struct PhiloxStreamV<float, PPT> {
  uint64_t group[PPT];
  float b[PPT][4];
  int pos;
  __device__ __forceinline__ void refill(const PhiloxKey& key, const PhiloxCtr& ctr) {
#pragma unroll
    for (int a = 0; a < PPT; ++a) {
      const uint4 w = philox_group_nowrap(ctr, key, group[a]);
      ++group[a];
      box_muller(w.x, w.y, &b[a][0], &b[a][1]);
      box_muller(w.z, w.w, &b[a][2], &b[a][3]);
    }
  }
  template <class Tab>
  __device__ __forceinline__ void init(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       const Tab&,
                                       const uint64_t (&first_element)[PPT]) {
#pragma unroll
    for (int a = 0; a < PPT; ++a) group[a] = philox_base_lo(ctr) + (first_element[a] >> 2);
    refill(key, ctr);
    pos = static_cast<int>(first_element[0] & 3);
  }
  template <class Tab>
  __device__ __forceinline__ void next(const PhiloxKey& key, const PhiloxCtr& ctr,
                                       const Tab&, float (&z)[PPT]) {
    if (pos == 4) {
      refill(key, ctr);
      pos = 0;
    }
#pragma unroll
    for (int a = 0; a < PPT; ++a)
      z[a] = pos == 0 ? b[a][0] : (pos == 1 ? b[a][1] : (pos == 2 ? b[a][2] : b[a][3]));
    ++pos;
  }
};

// Inverse-CDF transform of K Sobol integer points.
template <int K, class Tab>
__device__ __forceinline__ void sobol_normals(const Tab& tab, const uint32_t (&xb)[K],
                                              double (&z)[K], int = 0) {
  double t[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = sobol_centered_f64(xb[k]);
  fm::ndtri_t_v<K>(tab, t, z);
}
// The same with the constant high word of (2^21 + x32 2^-31) held in K registers
// the caller keeps alive (`hi[k]` = 0x41400000, opaque to the compiler): the
// integer point lands in the low half of a register pair whose high half is
// already in place, instead of costing one MOV per draw.
template <int K, class Tab>
__device__ __forceinline__ void sobol_normals(const Tab& tab, const uint32_t (&xb)[K],
                                              const int (&hi)[K], double (&z)[K], int = 0) {
  double t[K];
#pragma unroll
  for (int k = 0; k < K; ++k)
    t[k] = __hiloint2double(hi[k], static_cast<int>(xb[k])) - 2097153.0;
  fm::ndtri_t_v<K>(tab, t, z);
}

// N Sobol integer points -> standard normals, float32, the same arithmetic as
// fm::ndtri_t_f32 per draw but evaluated side by side: the central polynomial
// runs unconditionally for all N (independent Horner chains, no branch between
// them) and ONE rarely taken branch per batch patches the draws in the tails
// (|z| > 3.1, 0.2 % of them).
// `clamp` (documented non-reference mode, tqf_plan_set_sobol_clamp): beyond 2^24
// points RN(x) 2^-32 can be exactly 1.0 and the reference's own
// erfinv returns +inf; with clamp != 0 such a draw uses the largest float32
// below one instead (t = 1 - 2^-23).  Handled inside the rare tail branch: the
// strict hot path is unchanged.
template <int N>
__device__ __forceinline__ void sobol_normals_f32(const uint32_t (&xb)[N], float (&z)[N],
                                                  int clamp = 0) {
  float t[N], w[N], p[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    t[k] = fmaf(__uint2float_rn(xb[k]), 4.656612873077393e-10f, -1.0f);
    const float a = fmaf(-t[k], t[k], 1.0f);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(a));   // a is never subnormal
    w[k] = l2 * -0.693147182f;
  }
  const float cc[TQF_NDTRI_F32_C_N] = {TQF_NDTRI_F32_C_LIST};
  float y[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    y[k] = w[k] - TQF_NDTRI_F32_C_MID;
    p[k] = cc[0];
  }
#pragma unroll
  for (int i = 1; i < TQF_NDTRI_F32_C_N; ++i)
#pragma unroll
    for (int k = 0; k < N; ++k) p[k] = fmaf(p[k], y[k], cc[i]);
  float wmax = w[0];
#pragma unroll
  for (int k = 1; k < N; ++k) wmax = fmaxf(wmax, w[k]);
  if (!(wmax < 6.25f)) {
    const float ct[TQF_NDTRI_F32_T_N] = {TQF_NDTRI_F32_T_LIST};
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (!(w[k] < 6.25f)) {
        if (clamp && t[k] == 1.0f) {
          t[k] = 0.99999988079071044921875f;             // 2 (1 - 2^-24) - 1
          w[k] = -__logf(fmaf(-t[k], t[k], 1.0f));
        }
        const float yt = sqrtf(w[k]) - TQF_NDTRI_F32_T_MID;
        float q = ct[0];
#pragma unroll
        for (int i = 1; i < TQF_NDTRI_F32_T_N; ++i) q = fmaf(q, yt, ct[i]);
        p[k] = q;
      }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) z[k] = t[k] * p[k];
}

template <int K, class Tab>
__device__ __forceinline__ void sobol_normals(const Tab&, const uint32_t (&xb)[K],
                                              float (&z)[K], int clamp = 0) {
  // bit-identical to z[k] = ndtri(sobol_uniform_f32(xb[k])): (u - 0.5) 2 with
  // u = RN(x) 2^-32 equals RN(RN(x) 2^-31 - 1), the scaling by two commutes
  // with the rounding
  sobol_normals_f32<K>(xb, z, clamp);
}
template <int K, class Tab>
__device__ __forceinline__ void sobol_normals(const Tab& tab, const uint32_t (&xb)[K],
                                              const int (&)[K], float (&z)[K], int clamp = 0) {
  sobol_normals<K>(tab, xb, z, clamp);
}

template <typename Real, int DIM>
__device__ __forceinline__ Real select_component(const Real (&v)[DIM], int comp) {
  Real r = v[0];
#pragma unroll
  for (int j = 1; j < DIM; ++j) r = selp_real(v[j], r, j == comp ? 1 : 0);
  return r;
}

// ------------------------------------------------------------ payoffs -----
__device__ __forceinline__ double eval_payoff(const PayoffK& d, double x_final, double x_max,
                                              double x_min, double tangent = 0.0,
                                              double surv_up = 1.0, double surv_dn = 1.0) {
  double f = x_final, fmax = x_max, fmin = x_min;
  if (d.transform == TQF_TRANSFORM_EXP) {
    f = exp(f);
    fmax = exp(fmax);
    fmin = exp(fmin);
  }
  // d f / d state: f itself for the exponential transform
  const double fprime = d.transform == TQF_TRANSFORM_EXP ? f : 1.0;
  // A non-finite state (e.g. the +inf normal of a float32 Sobol uniform equal to
  // 1.0, SURVEY F7, or inf - inf = NaN one step later) makes the reference's
  // relu(...) NaN; `NaN > 0` is false, so without this test such a path would be
  // priced as 0 instead of being counted as non-finite by the caller.
  if (!isfinite(f)) return f - f;
  double v;
  switch (d.kind) {
    case TQF_PAYOFF_CALL_TANGENT:
      v = f - d.strike > 0.0 ? fprime * tangent : 0.0;
      break;
    case TQF_PAYOFF_PUT_TANGENT:
      v = d.strike - f > 0.0 ? -(fprime * tangent) : 0.0;
      break;
    case TQF_PAYOFF_CALL:
      v = f - d.strike > 0.0 ? f - d.strike : 0.0;
      break;
    case TQF_PAYOFF_PUT:
      v = d.strike - f > 0.0 ? d.strike - f : 0.0;
      break;
    // (with the bridge flag the payoff is weighted by the probability that the
    // continuous path between the grid points did not touch the barrier either)
    case TQF_PAYOFF_UP_OUT_CALL:
      v = (f - d.strike > 0.0 && !(fmax > d.barrier)) ? f - d.strike : 0.0;
      if (d.bridge) v *= surv_up;
      break;
    case TQF_PAYOFF_UP_OUT_PUT:
      v = (d.strike - f > 0.0 && !(fmax > d.barrier)) ? d.strike - f : 0.0;
      if (d.bridge) v *= surv_up;
      break;
    case TQF_PAYOFF_DOWN_OUT_PUT:
      v = (d.strike - f > 0.0 && !(fmin < d.barrier)) ? d.strike - f : 0.0;
      if (d.bridge) v *= surv_dn;
      break;
    case TQF_PAYOFF_DOWN_OUT_CALL:
      v = (f - d.strike > 0.0 && !(fmin < d.barrier)) ? f - d.strike : 0.0;
      if (d.bridge) v *= surv_dn;
      break;
    default:  // TQF_PAYOFF_IDENTITY
      v = f;
      break;
  }
  return v * d.scale;
}

// -------------------------------------------------------------- kernel ----
// Paths carried by one thread: enough that PPT * (draws per step) = 4 inverse
// CDFs (Sobol) or 4 Box-Muller pairs (Philox) are evaluated side by side.
// One step's NCOEF model constants through a generic pointer (shared or global
// table); 16-byte loads when the row size allows (rows start 16-byte aligned).
template <typename Real, int NCOEF>
__device__ __forceinline__ void load_step_coef(const Real* row, Real (&cc)[NCOEF]) {
  constexpr int PER16 = 16 / sizeof(Real);
  if ((NCOEF % PER16) == 0) {
    const uint4* v = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < NCOEF / PER16; ++i) {
      const uint4 q = v[i];
      if (sizeof(Real) == 8) {
        cc[2 * i] = static_cast<Real>(__hiloint2double(static_cast<int>(q.y), static_cast<int>(q.x)));
        cc[2 * i + 1] = static_cast<Real>(__hiloint2double(static_cast<int>(q.w), static_cast<int>(q.z)));
      } else {
        cc[4 * i] = static_cast<Real>(__uint_as_float(q.x));
        cc[4 * i + 1] = static_cast<Real>(__uint_as_float(q.y));
        cc[4 * i + 2] = static_cast<Real>(__uint_as_float(q.z));
        cc[4 * i + 3] = static_cast<Real>(__uint_as_float(q.w));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NCOEF; ++i) cc[i] = row[i];
  }
}

#ifndef TQF_SOBOL_DRAWS
#define TQF_SOBOL_DRAWS 8   // inverse CDFs evaluated side by side by one thread (tuning knob)
#endif
template <class Model, int RNGK>
struct PathsPerThread {
  static constexpr int value =
      (RNGK == RNGK_SOBOL) ? (Model::NF >= TQF_SOBOL_DRAWS ? 1 : TQF_SOBOL_DRAWS / Model::NF) : 4;
};

// PPT_: paths (antithetic pairs) carried by one thread.  The default suits runs that
// fill the machine; the Philox kernels also exist with PPT_ = 1 for small runs
// (config C1: 50 000 pairs), where four times as many CTAs -- ten warps per SM
// instead of less than three -- hide the latency of the dependent Philox /
// Box-Muller chain that a handful of warps per SM cannot.
template <class Model, int RNGK, bool ANTI, int MODE, int PPT_ = PathsPerThread<Model, RNGK>::value>
__global__ void __launch_bounds__(kBlock, TQF_MIN_BLOCKS)
path_kernel(const KParams<typename Model::Real> P) {
  using Real = typename Model::Real;
  constexpr int DIM = Model::DIM, NF = Model::NF, NCOEF = Model::NCOEF;
  constexpr int NPATH = ANTI ? 2 : 1;
  constexpr int PPT = PPT_;
  constexpr bool kPrice = MODE != MODE_PATHS;
  constexpr bool kExtrema = MODE == MODE_PRICE_EXTREMA || MODE == MODE_PRICE_BRIDGE;
  constexpr bool kBridge = MODE == MODE_PRICE_BRIDGE;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: log table (double Sobol only) | coef [S][NCOEF] Real | record_slot
  //         [S+1] int | sobol high [PPT][T] u32 | sobol low [T][8] u32 |
  //         accumulators [kWarps][8][3] double
  // {T, s} table of the table logarithm behind the FP64 inverse CDF: first, so
  // that its shared address is a compile-time constant in the LDS
  // (double Sobol: the ndtri table; double Philox: the MID-free table of the
  // Box-Muller logarithm, stored behind the first one in the device buffer)
  constexpr bool kLogTab = (RNGK == RNGK_SOBOL || RNGK == RNGK_PHILOX) && sizeof(Real) == 8;
  double* s_logtab = reinterpret_cast<double*>(smem_raw);
  size_t off = kLogTab ? static_cast<size_t>(TQF_LOGTAB_COUNT) * 2 * sizeof(double) : 0;
  Real* s_coef = reinterpret_cast<Real*>(smem_raw + off);
  if (P.tables_in_smem) {
    off += static_cast<size_t>(P.num_steps) * NCOEF * sizeof(Real);
    off = (off + 15) & ~static_cast<size_t>(15);
  }
  int* s_rec = reinterpret_cast<int*>(smem_raw + off);
  if (P.tables_in_smem)
    off += ((static_cast<size_t>(P.num_steps) + 1) * sizeof(int) + 15) & ~static_cast<size_t>(15);
  uint32_t* s_high = reinterpret_cast<uint32_t*>(smem_raw + off);
  if (RNGK == RNGK_SOBOL) off += static_cast<size_t>(PPT) * kSobolTileDims * sizeof(uint32_t);
  uint4* s_low = reinterpret_cast<uint4*>(smem_raw + off);
  if (RNGK == RNGK_SOBOL) off += static_cast<size_t>(kSobolTileDims) * 8 * sizeof(uint32_t);
  double* s_acc = reinterpret_cast<double*>(smem_raw + off);
  double* s_col = s_acc + kWarps * TQF_MAX_PAYOFFS * 3;   // [kWarps][colsum_cols]

  const int tid = threadIdx.x;
  if (kLogTab)
    fm::fill_smem_logtab(s_logtab, P.logtab + (RNGK == RNGK_PHILOX ? 2 * TQF_LOGTAB_COUNT : 0), tid,
                         kBlock);
  const fm::SmemTab tab = kLogTab ? fm::SmemTab(s_logtab) : fm::SmemTab();
  const Real* coef_tab = P.coef;
  const int* rec_tab = P.record_slot;
  if (P.tables_in_smem) {
    for (int i = tid; i < P.num_steps * NCOEF; i += kBlock) s_coef[i] = P.coef[i];
    coef_tab = s_coef;
    for (int i = tid; i <= P.num_steps; i += kBlock) s_rec[i] = P.record_slot[i];
    rec_tab = s_rec;
  }
  if (kPrice) {
    for (int i = tid; i < kWarps * TQF_MAX_PAYOFFS * 3; i += kBlock) s_acc[i] = 0.0;
  } else if (P.colsum_partials) {
    for (int i = tid; i < kWarps * P.colsum_cols; i += kBlock) s_col[i] = 0.0;
  }
  __syncthreads();

  // Sobol: masks of this thread's low index bits (the same for all its paths).
  uint32_t lowmask[kLowBits];
#pragma unroll
  for (int b = 0; b < kLowBits; ++b) {
    lowmask[b] = 0u - ((static_cast<uint32_t>(tid) >> b) & 1u);
    asm volatile("" : "+r"(lowmask[b]));  // keep in a register; do not rematerialise per draw
  }

  // high words of the Sobol -> double conversion (see sobol_normals)
  int t_hi[PPT * NF];
#pragma unroll
  for (int k = 0; k < PPT * NF; ++k) {
    t_hi[k] = (RNGK == RNGK_SOBOL && sizeof(Real) == 8) ? P.sobol_hi[k & 7] : 0x41400000;
  }

  constexpr int TILE_STEPS = (kSobolTileDims / NF) > 0 ? (kSobolTileDims / NF) : 1;
  const uint64_t stream_stride = static_cast<uint64_t>(P.num_steps_total) * NF;
  const uint64_t num_super = (P.num_chunks + PPT - 1) / PPT;

  for (uint64_t sc = blockIdx.x; sc < num_super; sc += gridDim.x) {
    // path a of this thread: Sobol index / unit number chunk_base + (sc*PPT+a)*128 + tid
    bool valid[PPT];
    uint64_t local[PPT];
    uint64_t first_element[PPT];
#pragma unroll
    for (int a = 0; a < PPT; ++a) {
      const uint64_t index = P.chunk_base + (sc * PPT + a) * kBlock + tid;
      valid[a] = index >= P.first_index && index < P.first_index + P.path_count;
      local[a] = index - P.first_index;  // row inside the shard
      first_element[a] =
          valid[a] ? ((P.path_offset + local[a]) * P.unit_stride + P.unit_offset) * stream_stride : 0;
    }

    Real x[PPT][NPATH][DIM], xmax[PPT][NPATH], xmin[PPT][NPATH];
    // bridge no-touch probabilities of the upper / lower barrier (MODE_PRICE_EXTREMA)
    double surv_up[PPT][NPATH], surv_dn[PPT][NPATH];
#pragma unroll
    for (int a = 0; a < PPT; ++a)
#pragma unroll
      for (int h = 0; h < NPATH; ++h) {
        surv_up[a][h] = 1.0;
        surv_dn[a][h] = 1.0;
      }
#pragma unroll
    for (int a = 0; a < PPT; ++a)
#pragma unroll
      for (int h = 0; h < NPATH; ++h)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
          Real x0j = P.x0[j];
          if (P.x0_paths != nullptr && valid[a])
            x0j = P.x0_paths[(P.path_offset + local[a] + h * P.x0_half) * DIM + j];
          x[a][h][j] = x0j;
          if (kExtrema && (j == 0 || j == P.monitor)) {
            xmax[a][h] = x0j;
            xmin[a][h] = x0j;
          }
        }

    // Stores the state of every path of this thread into time slot `slot`.
    auto store_slot = [&](int slot) {
      double cs[DIM];
#pragma unroll
      for (int j = 0; j < DIM; ++j) cs[j] = 0.0;
#pragma unroll
      for (int a = 0; a < PPT; ++a)
        if (valid[a]) {
#pragma unroll
          for (int h = 0; h < NPATH; ++h)
#pragma unroll
            for (int j = 0; j < DIM; ++j) {
              const Real v = P.store_exp ? static_cast<Real>(exp(x[a][h][j])) : x[a][h][j];
              P.out[static_cast<int64_t>(local[a] + h * P.anti_half) * P.stride_path +
                    slot * P.stride_time + j * P.stride_dim] = v;
              cs[j] += static_cast<double>(v);
            }
        }
      if (P.colsum_partials) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
          const double t = warp_sum(cs[j]);
          if (lane == 0) s_col[warp * P.colsum_cols + slot * DIM + j] += t;
        }
      }
    };
    if (MODE == MODE_PATHS) {
      const int slot = rec_tab[0];
      if (slot >= 0) store_slot(slot);
    }

    // Evaluates and reduces the payoffs attached to `step_index`.
    auto eval_payoffs = [&](int step_index) {
      const int warp = tid >> 5, lane = tid & 31;
      for (int q = 0; q < P.num_payoffs; ++q) {
        const PayoffK& d = P.pay[q];
        if (d.step != step_index) continue;
        double sum = 0.0, sq = 0.0, bad = 0.0;
#pragma unroll
        for (int a = 0; a < PPT; ++a) {
          if (valid[a]) {
#pragma unroll
            for (int h = 0; h < NPATH; ++h) {
              double v;
              if (d.kind == TQF_PAYOFF_HW_SWAPTION) {
                const SwaptionK& sw = P.swaptions[q];
                const double integral = static_cast<double>(x[a][h][DIM - 1]);
                double acc = 0.0;
                if (DIM <= 2 || sw.num_factors == 1) {
                  const double xs = static_cast<double>(x[a][h][0]);
                  for (int j = 0; j < sw.num_payments; ++j)
                    acc = fma(sw.coef[j], exp(fma(-sw.g[j], xs, sw.k[j])), acc);
                } else {
                  // several factors (TQF_MODEL_HJM): log P_j = k_j - sum_i g_ji x_i
                  const int nf = sw.num_factors;
                  for (int j = 0; j < sw.num_payments; ++j) {
                    double e = sw.k[j];
#pragma unroll
                    for (int i = 0; i < DIM - 1; ++i)
                      if (i < nf) e = fma(-sw.g[j * nf + i], static_cast<double>(x[a][h][i]), e);
                    acc = fma(sw.coef[j], exp(e), acc);
                  }
                }
                double swap = exp(-integral) * (1.0 - acc);
                swap = sw.is_payer ? swap : -swap;
                v = (swap > 0.0 ? swap : 0.0) * d.scale;
              } else {
                const double xf = static_cast<double>(select_component<Real, DIM>(x[a][h], d.component));
                const double xa = kExtrema ? static_cast<double>(xmax[a][h]) : xf;
                const double xi = kExtrema ? static_cast<double>(xmin[a][h]) : xf;
                const double tg = (d.kind == TQF_PAYOFF_CALL_TANGENT || d.kind == TQF_PAYOFF_PUT_TANGENT)
                                      ? static_cast<double>(select_component<Real, DIM>(x[a][h], d.tangent))
                                      : 0.0;
                v = eval_payoff(d, xf, xa, xi, tg, kBridge ? surv_up[a][h] : 1.0,
                                kBridge ? surv_dn[a][h] : 1.0);
              }
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
          double* acc = s_acc + (warp * TQF_MAX_PAYOFFS + q) * 3;
          acc[0] += sum;
          acc[1] += sq;
          acc[2] += bad;
        }
      }
    };
    if (kPrice && rec_tab[0] >= 0) eval_payoffs(0);

    // One Euler step of every path of this thread with the normals z, then the
    // running extrema and the payoffs / stores attached to the step.
    auto step_body = [&](int s, const Real (&z)[PPT][NF]) {
      const int rec_next = rec_tab[s + 1];
      Real cc[NCOEF];
      load_step_coef<Real, NCOEF>(coef_tab + static_cast<size_t>(s) * NCOEF, cc);
      Real xpre[PPT][NPATH], bvar[PPT][NPATH];
      if constexpr (HasBridgeVar<Model>::value && kBridge) {
        {
#pragma unroll
          for (int a = 0; a < PPT; ++a)
#pragma unroll
            for (int h = 0; h < NPATH; ++h) {
              xpre[a][h] = x[a][h][0];
              bvar[a][h] = Model::bridge_var(x[a][h], cc);
            }
        }
      }
      if constexpr (HasBatchStep<Model>::value) {
        Model::template step_batch<PPT, NPATH, 0>(x, z, cc, Real(1));
        if (ANTI) Model::template step_batch<PPT, NPATH, NPATH - 1>(x, z, cc, Real(-1));
      } else {
#pragma unroll
        for (int a = 0; a < PPT; ++a) {
          Model::step(x[a][0], z[a], cc);
          if (ANTI) {
            Real zm[NF];
#pragma unroll
            for (int j = 0; j < NF; ++j) zm[j] = -z[a][j];
            Model::step(x[a][NPATH - 1], zm, cc);
          }
        }
      }
      if constexpr (HasBridgeVar<Model>::value && kBridge) {
        {
          // brownian_bridge_single: P(no touch) = 1 - exp(-2 (x_s - b)(x_e - b) / var) when
          // both ends are on the inner side of the barrier, 0 otherwise
#pragma unroll
          for (int a = 0; a < PPT; ++a)
#pragma unroll
            for (int h = 0; h < NPATH; ++h) {
              const double xs = static_cast<double>(xpre[a][h]);
              const double xe = static_cast<double>(x[a][h][0]);
              const double var = static_cast<double>(bvar[a][h]);
              if (P.bridge & 1) {
                const double ds = P.bridge_up - xs, de = P.bridge_up - xe;
                const double p = (ds > 0.0 && de > 0.0)
                                     ? (var > 0.0 ? 1.0 - exp(-2.0 * (ds * de) / var) : 1.0) : 0.0;
                surv_up[a][h] *= p;
              }
              if (P.bridge & 2) {
                const double ds = xs - P.bridge_dn, de = xe - P.bridge_dn;
                const double p = (ds > 0.0 && de > 0.0)
                                     ? (var > 0.0 ? 1.0 - exp(-2.0 * (ds * de) / var) : 1.0) : 0.0;
                surv_dn[a][h] *= p;
              }
            }
        }
      }
      if (kPrice) {
        if (kExtrema) {
          // running extrema of the ONE monitored state component; the flags are
          // uniform, so each block is a branch around straight-line code
          if (DIM == 1 || P.monitor == 0) {
            if (P.need_extrema & 1) {
#pragma unroll
              for (int a = 0; a < PPT; ++a)
#pragma unroll
                for (int h = 0; h < NPATH; ++h)
                  xmax[a][h] = x[a][h][0] > xmax[a][h] ? x[a][h][0] : xmax[a][h];
            }
            if (P.need_extrema & 2) {
#pragma unroll
              for (int a = 0; a < PPT; ++a)
#pragma unroll
                for (int h = 0; h < NPATH; ++h)
                  xmin[a][h] = x[a][h][0] < xmin[a][h] ? x[a][h][0] : xmin[a][h];
            }
          } else {
#pragma unroll
            for (int a = 0; a < PPT; ++a)
#pragma unroll
              for (int h = 0; h < NPATH; ++h) {
                const Real xm = select_component<Real, DIM>(x[a][h], P.monitor);
                if (P.need_extrema & 1) xmax[a][h] = xm > xmax[a][h] ? xm : xmax[a][h];
                if (P.need_extrema & 2) xmin[a][h] = xm < xmin[a][h] ? xm : xmin[a][h];
              }
          }
        }
        if (rec_next >= 0) eval_payoffs(s + 1);
      } else {
        if (rec_next >= 0) store_slot(rec_next);
      }
    };

    // float64 Philox with one or two factors: a Philox group yields the normals of
    // two consecutive steps (NF = 1) or of the two factors of one step (NF = 2), so
    // the loop consumes whole groups and nothing selects between buffered halves.
    constexpr bool kPhiloxPairs = RNGK == RNGK_PHILOX && sizeof(Real) == 8 && NF <= 2;
    if (kPhiloxPairs) {
      uint64_t group[PPT];
#pragma unroll
      for (int a = 0; a < PPT; ++a) group[a] = philox_base_lo(P.ctr) + (first_element[a] >> 1);
      double g0[PPT], g1[PPT];
      Real z[PPT][NF];
      int s = 0;
      if (NF == 1) {
        // A path whose first element is odd (odd number of steps per path) starts in
        // the middle of a group -- same parity for all paths of a thread: its loop
        // starts at s = -1 and skips that half.  Two step bodies per group, each
        // guarded by a compare (the payoff code inside is instantiated twice, not
        // once per special case).
        s = (first_element[0] & 1) ? -1 : 0;
        for (; s < P.num_steps; s += 2) {
          philox_box_muller_v<PPT>(P.key, P.ctr, tab, group, g0, g1);
          if (s >= 0) {
#pragma unroll
            for (int a = 0; a < PPT; ++a) z[a][0] = static_cast<Real>(g0[a]);
            step_body(s, z);
          }
          if (s + 1 < P.num_steps) {
#pragma unroll
            for (int a = 0; a < PPT; ++a) z[a][0] = static_cast<Real>(g1[a]);
            step_body(s + 1, z);
          }
        }
      } else {
        for (; s < P.num_steps; ++s) {
          philox_box_muller_v<PPT>(P.key, P.ctr, tab, group, g0, g1);
#pragma unroll
          for (int a = 0; a < PPT; ++a) {
            z[a][0] = static_cast<Real>(g0[a]);
            z[a][NF - 1] = static_cast<Real>(g1[a]);
          }
          step_body(s, z);
        }
      }
      continue;                                   // next super-chunk
    }

    PhiloxStreamV<Real, PPT> stream;
    if (RNGK == RNGK_PHILOX) stream.init(P.key, P.ctr, tab, first_element);

    for (int s0 = 0; s0 < P.num_steps; s0 += TILE_STEPS) {
      const int s1 = min(P.num_steps, s0 + TILE_STEPS);
      if (RNGK == RNGK_SOBOL) {
        // Stage the direction numbers of dimensions [s0*NF, s1*NF): the XOR of
        // each chunk's common high index bits, and the kLowBits low columns.
        __syncthreads();
        for (int dd = tid; dd < (s1 - s0) * NF; dd += kBlock) {
          const uint32_t* v = P.sobol_v + (static_cast<size_t>(s0) * NF + dd) * 32;
          s_low[2 * dd] = *reinterpret_cast<const uint4*>(v);
          s_low[2 * dd + 1] = *reinterpret_cast<const uint4*>(v + 4);
#pragma unroll
          for (int a = 0; a < PPT; ++a) {
            uint32_t hb = static_cast<uint32_t>(
                (P.chunk_base + (sc * PPT + a) * kBlock) >> kLowBits);
            uint32_t h = 0;
            while (hb) {
              const int b = __ffs(hb) - 1;
              h ^= v[kLowBits + b];
              hb &= hb - 1;
            }
            s_high[a * kSobolTileDims + dd] = h;
          }
        }
        __syncthreads();
      }
      for (int s = s0; s < s1; ++s) {
        Real z[PPT][NF];
        if (RNGK == RNGK_PHILOX) {
#pragma unroll
          for (int j = 0; j < NF; ++j) {
            Real zz[PPT];
            stream.next(P.key, P.ctr, tab, zz);
#pragma unroll
            for (int a = 0; a < PPT; ++a) z[a][j] = zz[a];
          }
        } else if (RNGK == RNGK_SOBOL) {
          uint32_t xb[PPT * NF];
#pragma unroll
          for (int j = 0; j < NF; ++j) {
            const int dd = (s - s0) * NF + j;
            const uint4 l0 = s_low[2 * dd];
            const uint4 l1 = s_low[2 * dd + 1];
            uint32_t lowx = l0.x & lowmask[0];
            lowx ^= l0.y & lowmask[1];
            lowx ^= l0.z & lowmask[2];
            lowx ^= l0.w & lowmask[3];
            lowx ^= l1.x & lowmask[4];
            lowx ^= l1.y & lowmask[5];
            lowx ^= l1.z & lowmask[6];
#pragma unroll
            for (int a = 0; a < PPT; ++a) xb[a * NF + j] = lowx ^ s_high[a * kSobolTileDims + dd];
          }
          Real zz[PPT * NF];
          sobol_normals<PPT * NF>(tab, xb, t_hi, zz, P.sobol_clamp);
#pragma unroll
          for (int a = 0; a < PPT; ++a)
#pragma unroll
            for (int j = 0; j < NF; ++j) z[a][j] = zz[a * NF + j];
        } else {
#pragma unroll
          for (int a = 0; a < PPT; ++a)
#pragma unroll
            for (int j = 0; j < NF; ++j)
              z[a][j] = P.draws[first_element[a] + static_cast<size_t>(s) * NF + j];
        }
        step_body(s, z);
      }
    }

  }

  if (kPrice) {
    __syncthreads();
    for (int i = tid; i < TQF_MAX_PAYOFFS * 3; i += kBlock) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += s_acc[w * TQF_MAX_PAYOFFS * 3 + i];
      const int q = i / 3, k = i - q * 3;
      P.partials[(static_cast<size_t>(blockIdx.x) * TQF_MAX_PAYOFFS + q) * 4 + k] = v;
    }
  } else if (P.colsum_partials) {
    __syncthreads();
    for (int i = tid; i < P.colsum_cols; i += kBlock) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += s_col[w * P.colsum_cols + i];
      P.colsum_partials[static_cast<size_t>(blockIdx.x) * P.colsum_cols + i] = v;
    }
  }
}

// Deterministic final reduction of the per-CTA partials.
struct PeerK;
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int num_blocks,
                                       int num_payoffs, double* __restrict__ sums, const PeerK pk);

template <typename Real>
size_t path_kernel_smem(int ncoef, int num_steps, int rngk, int mode, bool tables_in_smem,
                        int ppt = kMaxPPT, int colsum_cols = 0) {
  size_t off = 0;
  if (tables_in_smem) {
    off = static_cast<size_t>(num_steps) * ncoef * sizeof(Real);
    off = (off + 15) & ~static_cast<size_t>(15);
    off += ((static_cast<size_t>(num_steps) + 1) * sizeof(int) + 15) & ~static_cast<size_t>(15);
  }
  if (rngk == RNGK_SOBOL) off += static_cast<size_t>(ppt) * kSobolTileDims * sizeof(uint32_t) + static_cast<size_t>(kSobolTileDims) * 8 * sizeof(uint32_t);
  off += static_cast<size_t>(kWarps) * TQF_MAX_PAYOFFS * 3 * sizeof(double);
  off += static_cast<size_t>(kWarps) * colsum_cols * sizeof(double);
  if ((rngk == RNGK_SOBOL || rngk == RNGK_PHILOX) && sizeof(Real) == 8)
    off += static_cast<size_t>(TQF_LOGTAB_COUNT) * 2 * sizeof(double);
  return off;
}

// A run whose chunks, grouped `ppt` per CTA, leave most of the machine empty (fewer
// CTAs than three per SM; max_grid = grid_per_sm() per SM): such runs take the one-path-per-
// thread Philox kernels.
// CTAs of the grid per SM (max_grid = that many per SM): 48 measured 0.3-1 % faster than 32 on
// C2 / C2-QE / C3 at full size and neutral at an eighth of it (more, shorter CTAs in the last
// wave); TQF_GRID_PER_SM overrides it for A/B runs.
inline int grid_per_sm() {
  static const int v = [] {
    const char* e = std::getenv("TQF_GRID_PER_SM");
    const int n = e ? std::atoi(e) : 0;
    return n > 0 ? n : 48;
  }();
  return v;
}
inline bool small_run(uint64_t num_chunks, int ppt, int max_grid) {
  return (num_chunks + ppt - 1) / ppt < static_cast<uint64_t>(max_grid / grid_per_sm() * 3);
}
// A Sobol run with fewer than eight waves of CTAs (three resident per SM): the last,
// partly filled wave would cost up to a sixth of the run (C2 sharded over 8 GPUs:
// 2442 chunks of 512 paths on 444 resident CTAs = 5.5 waves, measured efficiency
// 0.82).  Such runs carry half as many paths per thread -- twice the CTAs, half
// the quantum.
inline bool few_waves(uint64_t num_chunks, int ppt, int max_grid) {
  static const bool off = [] {
    const char* e = std::getenv("TQF_FEW_WAVES");      // "0": always the default kernels (A/B)
    return e && e[0] == '0';
  }();
  return !off && (num_chunks + ppt - 1) / ppt < static_cast<uint64_t>(max_grid / grid_per_sm() * 3) * 8;
}

// Models whose step is written for a full batch of paths per thread (the QE step compacts
// the lanes of its exponential branch across the batch): halving the batch costs them more
// than the shorter last wave returns (C2-QE at 1.25 M paths: 4.99 ms with the halved batch,
// 4.38 ms without; the Euler kernel: 2.04 ms either way, 2 % better halved at 625 k paths).
template <class M>
struct KeepsFullBatch : std::false_type {};
template <typename R>
struct KeepsFullBatch<HestonQeModel<R>> : std::true_type {};

// Launches the right instantiation for (rng kind, antithetic, mode).
template <class Model>
int launch_path_kernel(int rngk, bool anti, int mode, int max_grid, size_t smem_unused,
                       const KParams<typename Model::Real>& P, cudaStream_t stream,
                       int* grid_out) {
  (void)smem_unused;
#define TQF_LAUNCH_PPT(RK, AN, MD, PPTV)                                               \
  do {                                                                                 \
    auto kern = path_kernel<Model, RK, AN, MD, PPTV>;                                  \
    constexpr int ppt = PPTV;                                                          \
    const size_t smem = path_kernel_smem<typename Model::Real>(                        \
        Model::NCOEF, P.num_steps, RK, MD, P.tables_in_smem != 0, ppt,                 \
        P.colsum_partials ? P.colsum_cols : 0);                                        \
    const uint64_t num_super = (P.num_chunks + ppt - 1) / ppt;                         \
    if (smem > 48 * 1024)                                                              \
      TQF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       static_cast<int>(smem)));                       \
    /* Many more CTAs than are resident (max_grid = 48 per SM, ~16 waves): the    */   \
    /* CTAs of an SM drift out of phase, so the table staging of one overlaps the */   \
    /* arithmetic of the others (a grid of exactly the resident CTAs, which run   */   \
    /* in lock step, measured 8% slower on C2), and the last partial wave is      */   \
    /* short.  Each CTA walks its chunks with a grid stride: the per-CTA partial  */   \
    /* sums and their fixed-order reduction stay reproducible.                    */   \
    const int cap = max_grid;                                                          \
    int grid = static_cast<int>(num_super < static_cast<uint64_t>(cap)                 \
                                    ? num_super : static_cast<uint64_t>(cap));         \
    if (grid < 1) grid = 1;                                                            \
    *grid_out = grid;                                                                  \
    kern<<<grid, kBlock, smem, stream>>>(P);                                           \
    TQF_CUDA_OK(cudaGetLastError());                                                   \
    return TQF_OK;                                                                     \
  } while (0)
#define TQF_LAUNCH(RK, AN, MD)                                                         \
  do {                                                                                 \
    constexpr int dflt = PathsPerThread<Model, RK>::value;                             \
    if constexpr (RK == RNGK_PHILOX && dflt > 1) {                                     \
      if (small_run(P.num_chunks, dflt, max_grid)) TQF_LAUNCH_PPT(RK, AN, MD, 1);      \
    }                                                                                  \
    if constexpr (RK == RNGK_SOBOL && dflt >= 4 && !KeepsFullBatch<Model>::value) {    \
      if (few_waves(P.num_chunks, dflt, max_grid)) TQF_LAUNCH_PPT(RK, AN, MD, dflt / 2); \
    }                                                                                  \
    TQF_LAUNCH_PPT(RK, AN, MD, dflt);                                                  \
  } while (0)
  if (mode == MODE_PRICE) {
    if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PRICE);
    if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PRICE);
    if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PRICE);
    if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PRICE);
  } else if (mode == MODE_PRICE_EXTREMA) {
    if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PRICE_EXTREMA);
    if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PRICE_EXTREMA);
    if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PRICE_EXTREMA);
    if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PRICE_EXTREMA);
  } else if (mode == MODE_PRICE_BRIDGE) {
    if constexpr (HasBridgeVar<Model>::value) {
      if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PRICE_BRIDGE);
      if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PRICE_BRIDGE);
      if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PRICE_BRIDGE);
      if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PRICE_BRIDGE);
    }
  } else {
    if (rngk == RNGK_PHILOX && anti) TQF_LAUNCH(RNGK_PHILOX, true, MODE_PATHS);
    if (rngk == RNGK_PHILOX) TQF_LAUNCH(RNGK_PHILOX, false, MODE_PATHS);
    if (rngk == RNGK_SOBOL) TQF_LAUNCH(RNGK_SOBOL, false, MODE_PATHS);
    if (rngk == RNGK_DRAWS) TQF_LAUNCH(RNGK_DRAWS, false, MODE_PATHS);
  }
#undef TQF_LAUNCH
#undef TQF_LAUNCH_PPT
  set_error("unsupported rng / mode combination");
  return TQF_ERR_UNSUPPORTED;
}

// Launch descriptor of the multi-asset GBM kernel (tqf_mvgbm.cu).
struct MvLaunch {
  int dtype, dim, num_steps, num_steps_total, rngk, mode, max_grid;
  const void* coef_dev;      // Real [S][2]
  const double* x0;          // host [dim]
  const double* mu;          // host [dim]
  const double* sigma;       // host [dim]
  const double* chol;        // host [dim][dim] lower triangular
  PhiloxKey key;
  PhiloxCtr ctr;
  const uint32_t* sobol_v;
  const double* logtab;
  const float* ndtab;        // device_ndtri_f32_tab(): cubic table of the float32 inverse CDF
  const void* lsplit_dev;    // mvgbm_upload_split() table (dim > 8)
  uint64_t first_index, path_offset, path_count;
  int num_payoffs;
  const PayoffK* pay;
  double* partials;
  const int* record_dev;
  void* out;
  int64_t stride_path, stride_time, stride_dim;
  int store_exp;
  int exact_log;
  int sobol_clamp;
};

int launch_mvgbm(const MvLaunch& a, cudaStream_t stream, int* grid_out);
// Uploads the factor / mu / sigma in the order the split kernel (dim > 8) reads them.
int mvgbm_upload_split(const double* chol, const double* mu, const double* sigma, int dim,
                       int dtype, void** out_dev);

}  // namespace tqf
