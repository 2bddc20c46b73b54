/* tqf.h -- C ABI of the B200 Monte-Carlo path engine (libtqf.so).
 *
 * The reference (google/tf-quant-finance) has no FFI for this path: the Euler
 * sampler is Python calling TensorFlow ops.  This header is therefore the NEW
 * boundary that the Python mirror of the reference API (tff_b200) binds with
 * ctypes.  Each entry point cites the reference code whose device work it
 * replaces (paths relative to tf_quant_finance/ in the reference checkout).
 *
 * Conventions
 *   - every function returns an int status: TQF_OK (0) or a negative
 *     TQF_ERR_* code; `tqf_last_error()` returns a thread-local message;
 *   - `*_dev` pointers are CUDA device pointers owned by the caller (obtained
 *     zero-copy from framework tensors through DLPack on the Python side);
 *     every other pointer is host memory that is only read during the call;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream);
 *     all device work is enqueued on it and the call returns without
 *     synchronising unless stated otherwise;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with TQF_ERR_CUDA.
 */
#ifndef TQF_H_
#define TQF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQF_VERSION 100

/* status codes */
#define TQF_OK 0
#define TQF_ERR_INVALID_ARGUMENT (-1)
#define TQF_ERR_UNSUPPORTED (-2)
#define TQF_ERR_CUDA (-3)
#define TQF_ERR_IO (-4)

/* dtypes */
#define TQF_F32 0
#define TQF_F64 1

/* random number generators (math/random_ops/multivariate_normal.py:27-44) */
#define TQF_RNG_PHILOX 1 /* PSEUDO / STATELESS (+ *_ANTITHETIC)              */
#define TQF_RNG_SOBOL 2  /* SOBOL                                            */
#define TQF_RNG_DRAWS 3  /* caller-supplied normal_draws=                    */

/* models: the drift/volatility closures the Euler loop evaluates per step */
#define TQF_MODEL_AFFINE_1F 1     /* a = a0 + a1 x,  S = b0 + b1 x (dim 1)   */
#define TQF_MODEL_GBM_1F 2        /* a = mu x,  S = sigma x                  */
#define TQF_MODEL_HESTON_EULER 3  /* heston/heston_model.py:143-173          */
#define TQF_MODEL_HESTON_QE 4     /* heston/heston_model.py:322-572          */
#define TQF_MODEL_MVGBM 5         /* multivariate_geometric_brownian_motion  */
#define TQF_MODEL_LINEAR_1F 6     /* x' = A x + B + C z                      */
#define TQF_MODEL_HW1F 7          /* HW exact OU step + short-rate integral  */
#define TQF_MODEL_AFFINE_ND 8     /* a = a0 + A1 x, S = B, dim 2..4          */
/* TQF_MODEL_AFFINE_1F with the pathwise tangents carried alongside the path
 * (the forward-mode Jacobians of euler_sampling.py:393-402, 467-510 /
 * math/custom_loops.py:20-215): state = [x, dx/dx0, dx/dtheta], coef columns
 * dt, sqrt_dt, a0, a1, b0, b1, da0/dtheta, da1/dtheta, db0/dtheta, db1/dtheta;
 * x0 = [x0, 1, 0].                                                          */
#define TQF_MODEL_AFFINE_1F_TANGENT 9
/* Milstein scheme of the 1-d affine process (milstein_sampling.py:565-575):
 * coef columns dt, sqrt_dt, a0, a1, b0, b1; dS/dx = b1.                     */
#define TQF_MODEL_MILSTEIN_1F 10
/* Gaussian / quasi-Gaussian HJM with deterministic volatility
 * (hjm/gaussian_hjm.py:203-228, hjm/quasi_gaussian_hjm.py:232-289): F = dim - 1
 * Markov factors x_i and the path integral I of the short rate r = sum_i x_i +
 * f(0, t) that the models' discount factors are made of (gaussian_hjm.py:451-456
 * left-point rule, quasi_gaussian_hjm.py:486-490 right-point rule):
 *   x_i' = (x_i + dt (a0_i - k_i x_i)) + sum_j B_ij z_j sqrt_dt,
 *   I'   = I + cL sum_i x_i + cR sum_i x_i' + cF.
 * coef columns: dt, sqrt_dt, a0[F] (= sum_j y_ij, the deterministic state y of the
 * model), k[F], B[F][F] row-major (sqrt_rho_ij sigma_i), cL, cR, cF.  F = 1..3;
 * num_factors = the normals one step CONSUMES: F (Gaussian HJM) or F + F^2
 * (quasi-Gaussian: the reference simulates vec(y) next to x with zero volatility
 * rows, so its Wiener process has F + F^2 components of which the first F act);
 * supported (F, num_factors): (1,1) (1,2) (2,2) (2,6) (3,3).  x0 = 0.         */
#define TQF_MODEL_HJM 11
#define TQF_MODEL_HESTON_TANGENT 12 /* Heston Euler + tangents of (X, V) wrt one parameter: state [X, V, dX, dV] */

/* payoff kinds (reduced in-kernel; callers: e.g. hull_white/swaption.py:310) */
#define TQF_PAYOFF_CALL 1          /* max(f(X_T) - K, 0)                    */
#define TQF_PAYOFF_PUT 2           /* max(K - f(X_T), 0)                    */
#define TQF_PAYOFF_UP_OUT_CALL 3   /* call, knocked out if max_t f > B      */
#define TQF_PAYOFF_DOWN_OUT_PUT 4  /* put, knocked out if min_t f < B       */
#define TQF_PAYOFF_UP_OUT_PUT 5
#define TQF_PAYOFF_DOWN_OUT_CALL 6
#define TQF_PAYOFF_IDENTITY 7      /* f(X_T) (moments)                      */
#define TQF_PAYOFF_HW_SWAPTION 8   /* hull_white/swaption.py:291-312        */
/* pathwise derivative of the call / put payoff: 1{f > K} f'(X_T) T_T resp.
 * -1{K > f} f'(X_T) T_T with T = state[tangent_component]                  */
#define TQF_PAYOFF_CALL_TANGENT 9
#define TQF_PAYOFF_PUT_TANGENT 10

/* state transform applied before the payoff */
#define TQF_TRANSFORM_NONE 0
#define TQF_TRANSFORM_EXP 1 /* state is log-price */

#define TQF_MAX_PAYOFFS 8
#define TQF_MAX_SWAPTION_PAYMENTS 64

const char* tqf_last_error(void);
int tqf_version(void);
/* Number of visible CUDA devices (0 when there is none). */
int tqf_device_count(void);
/* sizeof of {tqf_rng_desc, tqf_model_desc, tqf_payoff_desc, tqf_lsm_desc} as
 * compiled into the library (lets a binding verify its struct layouts). */
int tqf_abi_sizes(int32_t out[4]);

/* ------------------------------------------------------------------------
 * Stand-alone generators (bit-exact stream tests; also `tff.math.random`).
 * ---------------------------------------------------------------------- */

/* TF `GenerateKey` (stateless seed scrambling) and `PhiloxRandom(seed,seed2)`
 * of a fresh stateful kernel: replaces the key/counter set-up inside
 * tf.random.stateless_normal / tf.random.normal
 * (math/random_ops/multivariate_normal.py:261-269).  Host only. */
int tqf_philox_stateless_key_counter(const int64_t seed[2], uint32_t key[2],
                                     uint32_t counter[4]);
int tqf_philox_stateful_key_counter(int64_t op_seed, uint32_t key[2],
                                    uint32_t counter[4]);

/* Raw Philox4x32-10 words of groups [first_group, first_group+num_groups):
 * out_dev is uint32 [num_groups][4]. */
int tqf_philox_raw_fill(const uint32_t key[2], const uint32_t counter[4],
                        uint64_t first_group, uint64_t num_groups,
                        uint32_t* out_dev, void* stream);

/* Elements [first_element, first_element+num_elements) of the flat normal
 * stream of tf.random.stateless_normal / tf.random.normal for `dtype`
 * (fp64: 2 elements per group, fp32: 4).  out_dev is dtype [num_elements]. */
int tqf_philox_normal_fill(const uint32_t key[2], const uint32_t counter[4],
                           uint64_t first_element, uint64_t num_elements,
                           int dtype, void* out_dev, void* stream);

/* The same for the uniform stream on [0, 1) of tf.random.stateless_uniform /
 * tf.random.uniform (math/random_ops/uniform.py:92-101): fp32 Uint32ToFloat,
 * four per group; fp64 Uint64ToDouble, two per group.                      */
int tqf_philox_uniform_fill(const uint32_t key[2], const uint32_t counter[4],
                            uint64_t first_element, uint64_t num_elements,
                            int dtype, void* out_dev, void* stream);

/* Non-randomized Halton points `halton.sample(dim, sequence_indices =
 * first_index .. first_index + count - 1, randomized=False)` in the reference's
 * floating-point arithmetic (math/random_ops/halton/halton_impl.py:250-288).
 *   radixes: host int32 [dim], the first `dim` primes (440-526);
 *   sizes:   host int32 [dim], digits kept per axis (_MAX_SIZES_BY_AXES, 530-534);
 *   weights: host double [dim][max_size], round(radix^j) in `dtype` (265-273);
 *   kind 1: uniforms, kind 2: normals sqrt(2) erfinv(2u - 1)
 *   (multivariate_normal.py:420).  out_dev is dtype [count][dim] row-major.
 * Synchronises `stream` before returning.                                   */
int tqf_halton_fill(const double* weights, const int32_t* sizes, const int32_t* radixes,
                    int dim, int max_size, uint64_t first_index, uint64_t count,
                    int kind, int dtype, void* out_dev, void* stream);

/* Owen-randomized Halton points (halton_impl.py:290-322, `randomized=True`).
 * tqf_halton_permutations builds, ON THE HOST, the digit permutations of
 * `_get_permutations` (342-379): for axis d (radix p_d) and digit position
 * i < num_coeffs the stable argsort of `tf.random.stateless_uniform([p_d],
 * seed=(seed + i, p_d), float64)` (math/random_ops/stateless.py:24-52);
 * perms is int32 [num_coeffs][sum_d p_d], axis d at column offset sum_{e<d} p_e
 * -- the flattened `perms` of the reference's HaltonParams.
 * tqf_halton_randomized_fill: as tqf_halton_fill with every digit looked up in
 * its permutation (perms_dev: the table above in DEVICE memory) and
 * zero_correction[d] (host double [dim], `stateless_uniform([dim,1], (seed,
 * seed), dtype) / p_d^size_d`, 303-322) added.  Synchronises `stream`.        */
int tqf_halton_permutations(int64_t seed, const int32_t* radixes, int dim, int num_coeffs,
                            int32_t* perms);
int tqf_halton_randomized_fill(const double* weights, const int32_t* sizes,
                               const int32_t* radixes, int dim, int max_size,
                               const int32_t* perms_dev, const double* zero_correction,
                               uint64_t first_index, uint64_t count, int kind, int dtype,
                               void* out_dev, void* stream);

/* Direction numbers m[dim][32] (int32) from the Joe-Kuo table: replaces
 * `load_data` + `_compute_direction_numbers`
 * (math/random_ops/sobol/sobol_impl.py:171-197, 237-261).
 *   poly_a[k]  = the column `a`, degree[k] = the column `s`,
 *   m_init     = [num_rows][18] initial m_i (zero padded).  Host only. */
int tqf_sobol_direction_numbers(const uint32_t* poly_a, const uint8_t* degree,
                                const uint32_t* m_init, int num_rows, int dim,
                                int32_t* out);
/* Same, parsing the original text file `new-joe-kuo-6.21201`. */
int tqf_sobol_direction_numbers_from_file(const char* path, int dim,
                                          int32_t* out);

/* Sobol points `sobol.sample(dim, num_results, skip)` in natural order
 * (sobol_impl.py:39-167).  direction_numbers: host int32 [dim][32].
 *   kind 0: int32 integer points x (num_digits bits)        -> int32 out
 *   kind 1: uniforms x / 2^num_digits in `dtype`            -> dtype out
 *   kind 2: normals sqrt(2) erfinv(2u-1)                    -> dtype out
 * (multivariate_normal.py:420).  out_dev is [num_results][dim] row-major.
 * `first_result` lets a shard produce rows [first_result, first_result+count)
 * of the num_results-row matrix (num_digits stays that of the full call). */
int tqf_sobol_fill(const int32_t* direction_numbers, int dim,
                   uint64_t num_results, uint64_t skip, uint64_t first_result,
                   uint64_t count, int kind, int dtype, void* out_dev,
                   void* stream);

/* ------------------------------------------------------------------------
 * tff.math.qmc: digital nets, Sobol generating matrices, lattice rules.
 * ---------------------------------------------------------------------- */

/* tf.random.stateless_uniform(shape, seed, minval, maxval, dtype=int32|int64)
 * as `_random_stateless_uniform` calls it (math/qmc/digital_net.py:155-202,
 * behind random_digital_shift 45-95 and random_scrambling_matrices 98-152):
 * `minval + x % (maxval - minval)` on the uint32 words (int_bits 32, four per
 * Philox call) or on uint64 pairs of words (int_bits 64).  out_dev is
 * int32 / int64 [num_elements].                                             */
int tqf_philox_uniform_int_fill(const uint32_t key[2], const uint32_t counter[4],
                                int64_t minval, int64_t maxval, uint64_t num_elements,
                                int int_bits, void* out_dev, void* stream);

/* `sobol_generating_matrices(dim, num_results, num_digits)`
 * (math/qmc/sobol.py:132-218, 221-243, 246-395) from the Joe-Kuo table in the
 * layout of tqf_sobol_direction_numbers; out is host int64
 * [dim][log_num_results], row 0 the identity.  Host only.                    */
int tqf_qmc_sobol_generating_matrices(const uint32_t* poly_a, const uint8_t* degree,
                                      const uint32_t* m_init, int num_rows, int dim,
                                      int log_num_results, int num_digits, int64_t* out);

/* `scramble_generating_matrices` (math/qmc/digital_net.py:422-527): column c of
 * coordinate d becomes XOR over the set bits (num_digits-1-shift) of G[d][c] of
 * S[d][shift] >> shift.  Host int64 [dim][num_columns] / [dim][scrambling_columns]
 * in, [dim][num_columns] out.  Host only.                                    */
int tqf_qmc_scramble_generating_matrices(const int64_t* generating_matrices,
                                         const int64_t* scrambling_matrices, int dim,
                                         int num_columns, int scrambling_columns, int num_digits,
                                         int64_t* out);

/* `digital_net_sample` (math/qmc/digital_net.py:205-419): point i, coordinate d
 * = real(shift[d] XOR_{bit b of index_i set, b < log_num_results} G[d][b]) /
 * real(1 << num_digits), optionally tent-transformed (utils.py:94-117).
 *   generating_matrices: host int64 [dim][num_columns] (already scrambled);
 *   digital_shift: host int64 [dim] or NULL;
 *   sequence_indices_dev: DEVICE int64 [count] or NULL for first_index + i;
 *   int_bits: 32 / 64 = the reference's int_dtype (integer wrap-around and the
 *   integer -> real cast follow it); dtype: TQF_F32 / TQF_F64.
 * out_dev is dtype [count][dim] row-major.  Synchronises `stream`.           */
int tqf_qmc_digital_net_fill(const int64_t* generating_matrices, int dim, int num_columns,
                             int log_num_results, const int64_t* digital_shift,
                             const int64_t* sequence_indices_dev, uint64_t first_index,
                             uint64_t count, int num_digits, int int_bits,
                             int apply_tent_transform, int dtype, void* out_dev, void* stream);

/* `lattice_rule_sample` (math/qmc/lattice_rule.py:99-229): floormod(real(i) *
 * floormod(z_d / n, 1) + shift_d, 1) with every product and sum rounded on its
 * own, as TensorFlow's element-wise ops are.  generating_vectors: host int64
 * [dim]; additive_shift: host double [dim] or NULL; indices as above.
 * Synchronises `stream`.                                                    */
int tqf_qmc_lattice_rule_fill(const int64_t* generating_vectors, int dim, int64_t num_results,
                              const double* additive_shift, const int64_t* sequence_indices_dev,
                              uint64_t first_index, uint64_t count, int int_bits,
                              int apply_tent_transform, int dtype, void* out_dev, void* stream);

/* ------------------------------------------------------------------------
 * The path engine: replaces the device work of
 *   models/euler_sampling.py:335-537  (_sample, _while_loop, _euler_step)
 *   models/utils.py:20-128            (generate_mc_normal_draws)
 * and of the payoff reductions of the callers.
 * ---------------------------------------------------------------------- */

typedef struct tqf_rng_desc {
  int32_t type;        /* TQF_RNG_*                                         */
  int32_t antithetic;  /* 1: path p >= N/2 uses -z of path p - N/2          */
  uint32_t key[2];     /* Philox key                                        */
  uint32_t counter[4]; /* Philox base counter                               */
  uint64_t skip;       /* Sobol: number of initial points skipped           */
  /* Sobol: host int32 [num_steps_total*num_factors][32] direction numbers  */
  const int32_t* direction_numbers;
  /* TQF_RNG_DRAWS: device pointer, dtype of the model, layout
   * [num_paths][num_steps_total][num_factors] (normal_draws= argument)     */
  const void* draws_dev;
  /* Draw unit of path p (batched calls, models/utils.py:98-107): the flat draw
   * index / Sobol point of path p is that of unit p * unit_stride + unit_offset
   * (unit_stride 0 is read as 1).  Sobol requires unit_stride == 1.         */
  uint64_t unit_stride;
  uint64_t unit_offset;
} tqf_rng_desc;

typedef struct tqf_model_desc {
  int32_t kind;            /* TQF_MODEL_*                                   */
  int32_t dtype;           /* TQF_F32 / TQF_F64                             */
  int32_t dim;             /* state dimension                               */
  int32_t num_factors;     /* normal draws per step                         */
  int32_t num_steps;       /* steps to execute                              */
  int32_t num_steps_total; /* len(all_times)-1: stride of the draw layout   */
  int32_t num_coef;        /* columns of `coef`                             */
  int32_t reserved;        /* MVGBM: 1 = exact log-space step (state = log x) */
  /* host double [num_steps][num_coef]; column meaning depends on `kind`
   * (documented in DESIGN.md); values are exactly representable in `dtype`.
   * Columns 0/1 are always dt and sqrt(dt).                                */
  const double* coef;
  const double* x0;        /* host double [dim] initial state               */
  const double* matrix;    /* MVGBM: host double [dim][dim] Cholesky factor */
  const double* vector;    /* MVGBM: host double [2][dim] means, vols       */
  /* Optional per-path initial states (`initial_state` of shape
   * [num_samples, dim], euler_sampling.py:357): DEVICE pointer, model dtype,
   * [num_paths_total][dim] row-major; NULL = every path starts at x0.  The
   * buffer must stay alive while the plan is used.  Not for TQF_MODEL_MVGBM. */
  const void* x0_paths_dev;
} tqf_model_desc;

typedef struct tqf_payoff_desc {
  int32_t kind;       /* TQF_PAYOFF_*                                       */
  int32_t component;  /* state component; -1 = arithmetic mean over dim     */
  int32_t transform;  /* TQF_TRANSFORM_*                                    */
  int32_t tangent_component; /* TQF_PAYOFF_*_TANGENT: state component of T   */
  double strike;
  double barrier;
  double scale;       /* multiplies the payoff (discount factor, notional)  */
  /* Every payoff is evaluated on the state after `expiry_step` steps
   * (0 = after the last step of the plan).                                 */
  int32_t expiry_step;
  /* TQF_PAYOFF_HW_SWAPTION (hjm/swaption_util.py:28-170 + swaption.py:300):
   *   P(t_e, T_j) = exp(pay_k[j] - pay_g[j] x),  x = r - f(0, t_e),
   *   payoff = scale max(+-exp(-I) (1 - sum_j pay_coef[j] P(t_e, T_j)), 0),
   * I = the path integral of the short rate carried by TQF_MODEL_HW1F.     */
  int32_t num_payments;
  int32_t is_payer;
  /* Barrier payoffs (TQF_PAYOFF_*_OUT_*): 1 = continuous monitoring by the
   * Brownian-bridge correction of black_scholes/brownian_bridge.py:118-196
   * (`brownian_bridge_single`): the payoff is multiplied by
   *   prod_steps [1 - exp(-2 (x_s - b)(x_e - b) / var_step)]
   * over the steps whose two ends lie on the inner side of the barrier b (0 as soon
   * as a grid point is beyond it), var_step = the variance of the monitored
   * component's increment given the state at the start of the step.  State
   * component 0 of TQF_MODEL_AFFINE_1F / LINEAR_1F / HESTON_EULER; with
   * TQF_TRANSFORM_EXP the barrier refers to exp(state) and the bridge runs on the
   * state (log-price) itself.                                               */
  int32_t brownian_bridge;
  /* TQF_PAYOFF_HW_SWAPTION on TQF_MODEL_HJM: number of factors F (0 is read as
   * 1); P(t_e, T_j) = exp(pay_k[j] - sum_i pay_g[j * F + i] x_i)
   * (quasi_gaussian_hjm.py:499-525), num_payments * F <= 64.                 */
  int32_t num_factors;
  int32_t reserved3;
  double reserved4;
  double pay_g[TQF_MAX_SWAPTION_PAYMENTS];     /* G(tau_j)=(1-e^{-k tau})/k */
  /* ln(P0(T_j)/P0(t_e)) - y(t_e) G_j^2 / 2 (vector_hull_white.py:783-814)  */
  double pay_k[TQF_MAX_SWAPTION_PAYMENTS];
  double pay_coef[TQF_MAX_SWAPTION_PAYMENTS];  /* coupon*tau_j (+1 on last) */
} tqf_payoff_desc;

typedef struct tqf_plan tqf_plan;

/* Builds a plan: validates the descriptors and uploads the small tables
 * (coefficients, direction numbers, Cholesky factor) to the current device.
 * `num_paths_total` is the GLOBAL number of paths N of the reference call
 * (it fixes the antithetic pairing and the Sobol num_digits); shards pass
 * their own [path_offset, path_offset+path_count) to the run calls. */
int tqf_plan_create(const tqf_model_desc* model, const tqf_rng_desc* rng,
                    uint64_t num_paths_total, tqf_plan** out_plan);
int tqf_plan_destroy(tqf_plan* plan);

/* Fused mode: simulate paths [path_offset, path_offset+path_count) and reduce
 * `num_payoffs` payoffs in-kernel.  sums_dev is double [num_payoffs][4]:
 * {sum, sum of squares, number of non-finite payoffs, unused}; UNNORMALISED so
 * that shards add (NCCL all-reduce across GPUs, then divide by N).
 * For antithetic plans path_offset/path_count address the first-half paths and
 * each unit contributes both partners.                                     */
int tqf_plan_price(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                   const tqf_payoff_desc* payoffs, int num_payoffs,
                   double* sums_dev, void* stream);

/* tqf_plan_price followed by the read-back of the sums into HOST memory
 * (sums_host: double [num_payoffs][4]) and a synchronisation of `stream`: the whole
 * device side of one pricing call of the reference API (`tf.reduce_mean(payoff(...))`
 * evaluated to a host value) behind a single FFI call.                       */
int tqf_plan_price_host(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                        const tqf_payoff_desc* payoffs, int num_payoffs,
                        double* sums_dev, double* sums_host, void* stream);

/* Materialising mode: writes the state at the recorded steps.
 *   record_slot: host int32 [num_steps+1]; entry 0 refers to the initial
 *   state, entry s+1 to the state after step s; value = output time slot or -1.
 *   out_dev[(p - path_offset)*stride_path + slot*stride_time + j*stride_dim]
 *   (strides in elements of the model dtype).  transform = TQF_TRANSFORM_EXP
 *   stores exp(state) (log-space models whose prices feed the LSM passes).  */
int tqf_plan_paths(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                   const int32_t* record_slot, void* out_dev,
                   int64_t stride_path, int64_t stride_time, int64_t stride_dim,
                   int transform, void* stream);

/* float32 Sobol draws beyond 2^24 points: the uniform RN(x) 2^-32 can be exactly
 * 1.0 and the reference's `erfinv((u - 0.5) * 2)` (multivariate_normal.py:420)
 * returns +inf (config C4 at 20 M paths).  Strict mode (default,
 * clamp = 0) reproduces that: the path's payoff is non-finite and is COUNTED in
 * sums[.][2] instead of summed.  clamp = 1 is a documented non-reference mode:
 * such a draw uses the largest float32 below one.  float64 is unaffected.   */
int tqf_plan_set_sobol_clamp(tqf_plan* plan, int clamp);

/* Multi-GPU pricing (one process per GPU of ONE box): once set, every
 * tqf_plan_price adds the sums of all ranks inside its reduction kernel, over
 * NVLink peer memory and in rank order (bit-identical sums on every rank), so
 * sums_dev comes back globally reduced -- replaces the ncclAllReduce of the
 * payoff sums.  bufs / epoch_base as tqf_lsm_set_peer_exchange; all
 * ranks must issue the same sequence of tqf_plan_price calls.               */
int tqf_plan_set_peer_exchange(tqf_plan* plan, int rank, int world, void* const* bufs,
                               uint64_t epoch_base);
int tqf_plan_peer_epoch(const tqf_plan* plan, uint64_t* epoch);

/* tqf_plan_paths that also returns, in column_sums_dev (double
 * [num_slots][dim], fully overwritten), the sum over this shard's paths of
 * every stored value (after `transform`): the basis-centring means of the
 * Longstaff-Schwartz passes (lsm.py:110-111) without a second pass over the
 * paths.  Slots that no step records receive 0.  Not for TQF_MODEL_MVGBM.  */
int tqf_plan_paths_sums(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                        const int32_t* record_slot, void* out_dev,
                        int64_t stride_path, int64_t stride_time, int64_t stride_dim,
                        int transform, int num_slots, double* column_sums_dev,
                        void* stream);

/* Hull-White discount curves along simulated short-rate paths: replaces the
 * device work of `sample_discount_curve_paths` / `_bond_reconstitution`
 * (models/hull_white/vector_hull_white.py:451-592, 783-814),
 *   out[n][i][j][d] = coef_a[i][j][d] exp(-(rates[n, j, d] - f0[j][d]) coef_g[i][j][d]),
 * with the path-independent factors folded into the tables by the caller:
 *   coef_g = (1 - e^{-a_d tau_i}) / a_d,
 *   coef_a = P0_d(t_j + tau_i) / P0_d(t_j) exp(-y_d(t_j) coef_g^2 / 2).
 *   rates_dev: model dtype, element (n, j, d) at n*rs_path + j*rs_time + d*rs_dim
 *     (strides in elements; the time-major view of tqf_plan_paths has rs_path = 1);
 *   f0_dev [k][dim], coef_a_dev / coef_g_dev [m][k][dim]: DEVICE doubles;
 *   out_dev: model dtype [num_paths][m][k][dim], contiguous.                 */
int tqf_hw_discount_curves(const void* rates_dev, int64_t rs_path, int64_t rs_time,
                           int64_t rs_dim, const double* f0_dev, const double* coef_a_dev,
                           const double* coef_g_dev, uint64_t num_paths, int m, int k, int dim,
                           int dtype, void* out_dev, void* stream);

/* HJM discount curves along simulated factor paths (`_bond_reconstitution`,
 * models/hjm/quasi_gaussian_hjm.py:499-525, behind `sample_discount_curve_paths`
 * 365-449):  out[n][i][j] = coef_a[i][j] exp(-sum_d coef_g[i][j][d] x[n, j, d]),
 *   coef_g = (1 - e^{-k_d tau_i}) / k_d,
 *   coef_a = P0(t_j + tau_i) / P0(t_j) exp(-coef_g' y(t_j) coef_g / 2)
 * (y is deterministic for deterministic volatility).  x_dev: model dtype, element
 * (n, j, d) at n*xs_path + j*xs_time + d*xs_dim; coef_a_dev [m][k], coef_g_dev
 * [m][k][num_factors]: DEVICE doubles; out_dev: model dtype [num_paths][m][k].
 * num_factors <= 3.                                                           */
int tqf_hjm_discount_curves(const void* x_dev, int64_t xs_path, int64_t xs_time,
                            int64_t xs_dim, const double* coef_a_dev,
                            const double* coef_g_dev, uint64_t num_paths, int m, int k,
                            int num_factors, int dtype, void* out_dev, void* stream);

/* Exercise values of Bermudan swaptions on Hull-White paths: replaces the bond
 * gather / weighted sum / scatter of hull_white/swaption.py:608-724
 * (`_map_payoff_to_sim_times`):
 *   values[u][n][b] = relu(1 - sum_j coef[b][e][j] exp(k[b][e][j] - g[b][e][j] x[n][u])),
 *   u = ex_slot[b][e] (slot of the e-th exercise date of swaption b among the unique
 *   exercise dates); x = r - f(0, t) at those dates, element (n, u) at
 *   n*xs_path + u*xs_slot; g / k / coef: DEVICE doubles [B][E][m] (coef carries the
 *   +1 of the float leg on its last entry); values_dev: model dtype [U][N][B],
 *   zero-filled by the caller (dates without exercise stay 0).               */
int tqf_hw_exercise_values(const void* x_dev, int64_t xs_path, int64_t xs_slot,
                           const double* g_dev, const double* k_dev, const double* coef_dev,
                           const int32_t* ex_slot_dev, uint64_t num_paths, int num_swaptions,
                           int num_exercise, int num_payments, int dtype, void* values_dev,
                           void* stream);

/* ------------------------------------------------------------------------
 * Longstaff-Schwartz regression passes on materialised paths: replaces the
 * device work of models/longstaff_schwartz/lsm.py:231-436 (payoff_fn, basis_fn,
 * the masked X'X / X'y matmuls, tf.where updates) for
 *   payoff_fn = make_basket_put_payoff(strikes)   (payoff_utils.py:27-97)
 *   basis_fn  = make_polynomial_basis(degree)      (lsm.py:50-125).
 * The K x K pseudo-inverse is solved by the caller between passes (after the
 * all-reduce of the sums when paths are sharded over GPUs).
 * ---------------------------------------------------------------------- */
typedef struct tqf_lsm_desc {
  int32_t dtype;       /* TQF_F32 / TQF_F64 (dtype of the paths)             */
  int32_t dim;         /* state dimension (<= 8)                             */
  int32_t batch;       /* number of payoffs B (strikes)                      */
  int32_t basis_size;  /* K = (degree+1)^dim (<= 128)                        */
  const int32_t* exponents; /* host [K][dim] monomial exponents              */
  const double* strikes;    /* host [B]                                      */
  uint64_t num_paths;       /* local paths                                   */
  uint64_t path_offset;     /* global index of local path 0                  */
  uint64_t num_calibration_samples; /* regress on global paths < n; 0 = all  */
  const void* paths_dev;    /* element (n, t, j) at n*stride_path +          */
  int64_t stride_path;      /*   t*stride_time + j*stride_dim (+ b*stride_   */
  int64_t stride_time;      /*   batch for batched sample paths), strides in */
  int64_t stride_dim;       /*   elements                                    */
  int64_t stride_batch;
  /* Optional caller-owned workspaces (e.g. from the framework's caching
   * allocator; NULL = the library allocates): W [B][num_paths] in `dtype`, and
   * partial sums of `partials_doubles` doubles (tqf_lsm_workspace). */
  void* w_dev;
  double* partials_dev;
  uint64_t partials_doubles;
  /* Optional: tabulated exercise values and per-path discounting (Bermudan
   * swaptions on short-rate paths, hull_white/swaption.py:608-724, where the
   * reference passes a payoff closure over precomputed swap values and rank-3
   * discount factors).  With `exercise_time_indices` (host, [num_exercise_times],
   * the time index of every exercise date in order) set:
   *   exercise_values_dev: `dtype` [num_exercise_times][B][num_paths], the value
   *     of exercising payoff b on path n at date t; replaces relu(strike - mean x);
   *   path_ratio_dev: `dtype` [num_exercise_times][num_paths], entry [e][n] =
   *     df[e+1]/df[e] of path n (lsm.py:304-325); replaces the ratio_* arguments
   *     of tqf_lsm_step, and entry [0] weights tqf_lsm_value_sum.
   * Either pointer may be NULL.  K <= 6 only. */
  const int32_t* exercise_time_indices;
  int32_t num_exercise_times;
  int32_t reserved;
  const void* exercise_values_dev;
  const void* path_ratio_dev;
} tqf_lsm_desc;

typedef struct tqf_lsm tqf_lsm;

/* Doubles of partial-sum workspace a handle for this descriptor can need
 * (`num_times` = number of exercise dates passed to tqf_lsm_column_sums). */
int tqf_lsm_workspace(const tqf_lsm_desc* desc, int num_times, uint64_t* partials_doubles);
int tqf_lsm_create(const tqf_lsm_desc* desc, tqf_lsm** out);
int tqf_lsm_destroy(tqf_lsm* lsm);
/* sums_dev[b][t][j] = sum_n x[n, time_indices[t], j] (basis centring). */
int tqf_lsm_column_sums(tqf_lsm* lsm, const int32_t* time_indices, int num_times,
                        double* sums_dev, void* stream);
/* W[b][n] = payoff_b(x[n, time_index, :]) (the terminal cashflow). */
int tqf_lsm_init(tqf_lsm* lsm, int time_index, void* stream);
/* One pass: optionally W' = ev > relu(X beta) ? ev : ratio_update W at time
 * index t_update, then optionally accumulate the normal equations at t_acc with
 * y = ratio_acc W'.  DEVICE arrays: mean_* (payoff b at b*mean_stride, `dim`
 * entries), beta [B][K], ratio_* [B].  sums_dev: double [B][num_sums]
 * (layout: tqf_lsm_sums_layout); NULL leaves the per-CTA partials for a fused
 * tqf_lsm_solve(reduce_partials = 1).  Nothing synchronises with the host. */
int tqf_lsm_step(tqf_lsm* lsm, int do_update, int t_update, const double* mean_update_dev,
                 const double* beta_dev, const double* ratio_update_dev, int do_accumulate,
                 int t_acc, const double* mean_acc_dev, const double* ratio_acc_dev,
                 int64_t mean_stride, double* sums_dev, void* stream);
/* beta_dev[b] = pinv(X'X_b) X'y_b (Jacobi eigen-solver on the device, singular
 * values below rcond * max dropped like tf.linalg.pinv).  reduce_partials = 1:
 * the preceding tqf_lsm_step was called with sums_dev = NULL and this call first
 * reduces its per-CTA partials into sums_dev (single-GPU fast path, one launch
 * less per date); 0: sums_dev already holds the (all-reduced) sums.  Packed
 * layout (K <= 6) only; TQF_ERR_UNSUPPORTED otherwise (solve on the host). */
int tqf_lsm_solve(tqf_lsm* lsm, double* sums_dev, int reduce_partials, double rcond,
                  double* beta_dev, void* stream);
/* Fused regression solve -- `beta = pinv(X'X) X'y` of lsm.py:369-377, which the
 * reference evaluates as separate matmul / pinv ops per exercise date -- for
 * the single-asset vectorised pass (dim 1,
 * K <= 6, contiguous time-major paths, even path count): once set, every
 * tqf_lsm_step that accumulates also reduces its per-CTA partials into
 * sums_dev [B][27] and writes beta [B][K] for the accumulated date, from the
 * last CTA to finish; the tqf_lsm_solve call that follows (same beta_dev)
 * then returns without launching anything.  Steps served by the other kernels
 * are unaffected.  ticket_dev: one zero-initialised uint32 in device memory.
 * Single GPU only (the sums are not all-reduced).                          */
int tqf_lsm_set_fused_solve(tqf_lsm* h, double rcond, double* sums_dev, double* beta_dev,
                            uint32_t* ticket_dev);

/* 1 when the fused pass (tqf_lsm_set_fused_solve) applies to this handle. */
int tqf_lsm_fused_eligible(const tqf_lsm* h, int* eligible);

/* The whole backward induction of lsm.py:296-330 after tqf_lsm_init, for a
 * handle with the fused solve set: one accumulate-only pass for the last
 * exercise date, then one update + accumulate pass per earlier date, launched
 * back to back from native code (the per-date host work of a Python loop,
 * ~50 us, is longer than a pass at a few million paths).
 *   exercise_times: host int32 [num_times] time indices;
 *   means_dev: device double, the basis means of exercise index e at
 *     means_dev + (e - 1) * dim (+ payoff * mean_stride);
 *   ratio_dev: device double [num_times][batch], row e = df[e + 1] / df[e].  */
int tqf_lsm_run_fused(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                      const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                      double* beta_dev, void* stream);

/* The whole backward induction of lsm.py:231-330 -- terminal cashflow
 * (lsm.py:266-276), one regression + exercise sweep per earlier date
 * (`_lsm_loop_body`, 403-436) and the sum of the discounted values (284-291) --
 * in ONE persistent cooperative launch: one CTA per SM owns a fixed set of path
 * tiles, streams the two path columns of every date through a TMA-fed
 * shared-memory ring, keeps W in L2, and the per-date regression is a grid
 * barrier whose last arriver reduces (and, with tqf_lsm_set_peer_exchange,
 * exchanges over NVLink), solves and releases.  Applies when
 * tqf_lsm_persistent_eligible (one payoff, dim 1, K <= 6, contiguous time-major
 * paths, 16-byte aligned columns, path count a multiple of 16 / sizeof(dtype)).
 *   means_dev / ratio_dev: as tqf_lsm_run_fused;
 *   skip_below: global paths below it do not count in the value sum
 *     (num_calibration_samples, lsm.py:286-289);
 *   value_sums_dev: double [2] = {sum of W over the counted paths, their count}
 *     -- over ALL ranks when the peer exchange is set;
 *   beta_dev: double [K] scratch (regression coefficients of the current date);
 *   history_dev: NULL, or double [num_times - 1][27 + 6]: row j = the reduced
 *     normal equations (tqf_lsm_sums_layout, packed) and beta of exercise index
 *     num_times - 1 - j (diagnostics / parity tests).
 * Nothing synchronises with the host; tqf_lsm_status reads back (and waits for)
 * a non-zero code when a grid barrier or peer timed out.                    */
int tqf_lsm_persistent_eligible(const tqf_lsm* h, int* eligible);
int tqf_lsm_run_persistent(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                           const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                           double rcond, uint64_t skip_below, double* value_sums_dev,
                           double* beta_dev, double* history_dev, void* stream);
int tqf_lsm_status(const tqf_lsm* h, uint64_t* status);

/* Multi-GPU (one process per GPU of ONE box): the reduced normal equations of
 * every exercise date are summed over the ranks INSIDE the tail of the fused
 * pass, by peer stores / flags over NVLink -- replaces the per-date
 * ncclAllReduce of K^2 + K doubles (lsm.py:369-377 computes lhs and
 * rhs per date).  Every rank adds the contributions in rank order, so all ranks
 * solve from bit-identical sums.
 *   bufs[r]: the exchange buffer of rank r (tqf_lsm_peer_bytes() bytes, zeroed
 *     once, allocated with tqf_peer_alloc by its owner and mapped here with
 *     tqf_peer_open), r = 0 .. world-1, world <= 8, batch <= 16;
 *   epoch_base: number of exchanges already performed on these buffers -- the
 *     same on every rank; read it back with tqf_lsm_peer_epoch after the call.
 * Requires tqf_lsm_fused_eligible (and, for the one-launch-per-date route,
 * tqf_lsm_set_fused_solve) on EVERY rank: all ranks must take the same route.  A peer that does not arrive within
 * ~10 s poisons the sums with NaN instead of hanging the device.           */
int tqf_lsm_peer_bytes(uint64_t* bytes);
int tqf_lsm_set_peer_exchange(tqf_lsm* h, int rank, int world, void* const* bufs,
                              uint64_t epoch_base);
int tqf_lsm_peer_epoch(const tqf_lsm* h, uint64_t* epoch);

/* Peer-visible device memory (CUDA IPC): alloc + export on the owner, open /
 * close on the other processes of the box, free on the owner.              */
int tqf_peer_alloc(uint64_t bytes, void** dev_ptr, uint8_t ipc_handle[64]);
int tqf_peer_open(const uint8_t ipc_handle[64], void** dev_ptr);
int tqf_peer_close(void* dev_ptr);
int tqf_peer_free(void* dev_ptr);
/* Status words of a rank's OWN exchange buffer: how many in-kernel exchanges
 * gave up waiting for a peer (their sums were poisoned with NaN) and the epoch
 * of the last one.  Synchronous 16-byte copy; the host calls it when a result
 * that went through an exchange is not finite.                              */
int tqf_peer_status(const void* own_buf, uint64_t* timeouts, uint64_t* last_epoch);


/* num_sums doubles per payoff.  Packed (K <= 6): the upper triangle of a 6 x 6
 * X'X row by row (21 entries) followed by 6 entries of X'y; otherwise X'X
 * [K][K] row-major followed by X'y [K]. */
int tqf_lsm_sums_layout(const tqf_lsm* lsm, int* num_sums, int* is_packed_symmetric);
/* sums_dev[b] = {sum_n W[b][n], count} over global paths >= skip_below. */
int tqf_lsm_value_sum(tqf_lsm* lsm, uint64_t skip_below, double* sums_dev, void* stream);

/* Measures the FP64 DFMA issue peak of the current device (Ginstr/s): the
 * roofline denominator for the fused mode (not in MEASURED_PEAKS.json).
 * Synchronises. */
int tqf_measure_fp64_peak(double* dfma_per_second, double* ffma_per_second);

/* Test hook: evaluates one of the hand-written device math functions
 * (csrc/tqf_math.cuh) elementwise on a device array, so that their accuracy
 * can be checked against a multiprecision reference.
 *   fn 0: log(x)  1: sqrt(x)  2: ndtri(x)  3: sin(x)  4: cos(x)  (x in the
 *   domains stated in tqf_math.cuh). */
int tqf_math_eval(int fn, const double* in_dev, double* out_dev, uint64_t n,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TQF_H_ */
