// Host side of the path engine: plans (device-resident tables), the fused
// pricing launch and the path-materialising launch.  C ABI in include/tqf.h.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "tqf_paths_kernel.cuh"
#include "tqf_peer.cuh"

namespace tqf {

int upload_sobol_table(const int32_t* direction_numbers, int dim, uint32_t** out_dev,
                       cudaStream_t stream);

#define TQF_EXTERN_MODEL(M)                                                             \
  extern template int launch_path_kernel<M<double>>(int, bool, int, int, size_t,        \
                                                    const KParams<double>&, cudaStream_t, int*); \
  extern template int launch_path_kernel<M<float>>(int, bool, int, int, size_t,         \
                                                   const KParams<float>&, cudaStream_t, int*);
TQF_EXTERN_MODEL(AffineModel1F)
TQF_EXTERN_MODEL(GbmModel1F)
TQF_EXTERN_MODEL(LinearModel1F)
TQF_EXTERN_MODEL(HestonEulerModel)
TQF_EXTERN_MODEL(HullWhite1FModel)
TQF_EXTERN_MODEL(HestonQeModel)
TQF_EXTERN_MODEL(AffineModel2D)
TQF_EXTERN_MODEL(AffineModel3D)
TQF_EXTERN_MODEL(AffineModel4D)
TQF_EXTERN_MODEL(HjmModel11)
TQF_EXTERN_MODEL(HjmModel12)
TQF_EXTERN_MODEL(HjmModel22)
TQF_EXTERN_MODEL(HjmModel26)
TQF_EXTERN_MODEL(HjmModel33)
#undef TQF_EXTERN_MODEL

__global__ void reduce_partials_kernel(const double* __restrict__ partials, int num_blocks,
                                       int num_payoffs, double* __restrict__ sums, const PeerK pk) {
  // Thread t sums the rows t, t + 1024, ... (independent loads: at most 5 rounds of
  // L2 latency for the 4736-CTA grid), then every (payoff, statistic) is combined by a
  // shuffle tree and across the 32 warps in a fixed order -> reproducible for a given
  // grid.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  __shared__ double s_w[32][TQF_MAX_PAYOFFS * 3];
  double acc[TQF_MAX_PAYOFFS * 3];
#pragma unroll
  for (int i = 0; i < TQF_MAX_PAYOFFS * 3; ++i) acc[i] = 0.0;
#pragma unroll 2
  for (int b = threadIdx.x; b < num_blocks; b += blockDim.x) {
    const double* row = partials + static_cast<size_t>(b) * TQF_MAX_PAYOFFS * 4;
#pragma unroll
    for (int q = 0; q < TQF_MAX_PAYOFFS; ++q)
      if (q < num_payoffs) {
        const double4 v = *reinterpret_cast<const double4*>(row + q * 4);
        acc[q * 3] += v.x;
        acc[q * 3 + 1] += v.y;
        acc[q * 3 + 2] += v.z;
      }
  }
#pragma unroll
  for (int i = 0; i < TQF_MAX_PAYOFFS * 3; ++i)
    if (i < num_payoffs * 3) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_w[warp][i] = v;
    }
  __syncthreads();
  if (threadIdx.x < num_payoffs * 4) {
    const int q = threadIdx.x >> 2, k = threadIdx.x & 3;
    double v = 0.0;
    if (k < 3)
      for (int w = 0; w < nwarps; ++w) v += s_w[w][q * 3 + k];
    sums[q * 4 + k] = v;
  }
  // several GPUs: the sums of all ranks are added here, over NVLink peer memory
  // and in rank order (bit-identical on every rank) -- no NCCL call per pricing
  if (pk.peer_world > 1) {
    __syncthreads();
    peer_all_reduce(pk, sums, num_payoffs * 4);
  }
}

struct ModelInfo {
  int dim, nf, ncoef;
};

static bool model_info(int kind, int dim, int num_factors, ModelInfo* info) {
  switch (kind) {
    case TQF_MODEL_HJM: {
      const int f = dim - 1;
      *info = {dim, num_factors, 5 + 2 * f + f * f};
      return (f == 1 && (num_factors == 1 || num_factors == 2)) ||
             (f == 2 && (num_factors == 2 || num_factors == 6)) || (f == 3 && num_factors == 3);
    }
    case TQF_MODEL_MVGBM: *info = {dim, dim, 2}; return dim >= 1 && dim <= 64;
    case TQF_MODEL_AFFINE_1F: *info = {1, 1, 6}; return true;
    case TQF_MODEL_AFFINE_1F_TANGENT: *info = {3, 1, 10}; return true;
    case TQF_MODEL_HESTON_TANGENT: *info = {4, 2, 12}; return true;
    case TQF_MODEL_MILSTEIN_1F: *info = {1, 1, 6}; return true;
    case TQF_MODEL_AFFINE_ND:
      *info = {dim, dim, 2 + dim + 2 * dim * dim};
      return dim >= 2 && dim <= 4;
    case TQF_MODEL_GBM_1F: *info = {1, 1, 4}; return true;
    case TQF_MODEL_LINEAR_1F: *info = {1, 1, 5}; return true;
    case TQF_MODEL_HESTON_EULER: *info = {2, 2, 6}; return true;
    case TQF_MODEL_HW1F: *info = {2, 1, 5}; return true;
    case TQF_MODEL_HESTON_QE: *info = {2, 2, 10}; return true;
    default: return false;
  }
}

}  // namespace tqf

using namespace tqf;

struct tqf_plan {
  tqf_model_desc model;
  tqf_rng_desc rng;
  ModelInfo info;
  uint64_t num_paths_total;
  int device;
  int max_grid;
  void* coef_dev;           // Real [num_steps][ncoef]
  uint32_t* sobol_dev;      // [S_total*nf][32]
  const double* logtab_dev; // shared per-device log table (not owned)
  const float* ndtab_dev;   // shared per-device float32 ndtri table (not owned; MVGBM)
  void* lsplit_dev;         // MVGBM dim > 8: factor in the split kernel's order
  PeerHost peer;            // tqf_plan_set_peer_exchange (world <= 1: single GPU)
  int sobol_clamp;          // tqf_plan_set_sobol_clamp
  double* colsum_dev;       // [max_grid][slots * dim] column-sum partials (lazily allocated)
  size_t colsum_doubles;
  double* partials_dev;     // [max_grid][TQF_MAX_PAYOFFS][4]
  int* record_dev;          // [num_steps+1]
  // what record_dev holds (and the stream it was uploaded on): a pricing loop
  // re-sends identical flags every call, a pageable 1 KB copy in front of the kernel
  std::vector<int>* record_cache;
  cudaStream_t record_stream;
  SwaptionK* swaptions_dev; // [TQF_MAX_PAYOFFS]
  double x0[64];
  double mu[64], sigma[64];
  double chol[64 * 64];
};

template <typename Real>
static int upload_coef(const tqf_model_desc* m, int ncoef, void** out) {
  const size_t n = static_cast<size_t>(m->num_steps) * ncoef;
  std::vector<Real> host(n > 0 ? n : 1);
  for (size_t i = 0; i < n; ++i) host[i] = static_cast<Real>(m->coef[i]);
  void* dev = nullptr;
  const int rc = dev_alloc(&dev, host.size() * sizeof(Real));
  if (rc != TQF_OK) return rc;
  cudaError_t e = cudaMemcpy(dev, host.data(), host.size() * sizeof(Real), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    dev_release(&dev, 1);
    return cuda_fail(e, "upload_coef");
  }
  *out = dev;
  return TQF_OK;
}

template <typename Real>
static void fill_common(const tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                        KParams<Real>* P) {
  std::memset(P, 0, sizeof(*P));
  P->coef = static_cast<const Real*>(plan->coef_dev);
  P->num_steps = plan->model.num_steps;
  P->num_steps_total = plan->model.num_steps_total;
  for (int j = 0; j < 4; ++j) P->x0[j] = static_cast<Real>(plan->x0[j]);
  P->x0_paths = static_cast<const Real*>(plan->model.x0_paths_dev);
  P->x0_half = plan->num_paths_total / 2;
  P->key = PhiloxKey{plan->rng.key[0], plan->rng.key[1]};
  P->ctr = PhiloxCtr{plan->rng.counter[0], plan->rng.counter[1], plan->rng.counter[2],
                     plan->rng.counter[3]};
  P->sobol_v = plan->sobol_dev;
  P->sobol_clamp = plan->sobol_clamp;
  P->logtab = plan->logtab_dev;
  for (int k = 0; k < 8; ++k) P->sobol_hi[k] = 0x41400000;
  P->draws = static_cast<const Real*>(plan->rng.draws_dev);
  P->path_offset = path_offset;
  P->path_count = path_count;
  P->unit_stride = plan->rng.unit_stride ? plan->rng.unit_stride : 1;
  P->unit_offset = plan->rng.unit_offset;
  P->first_index = plan->rng.type == TQF_RNG_SOBOL
                       ? plan->rng.skip + 1 + plan->rng.unit_offset + path_offset
                       : path_offset;
  P->chunk_base = P->first_index & ~static_cast<uint64_t>(kBlock - 1);
  const uint64_t end = P->first_index + path_count;
  P->num_chunks = (end - P->chunk_base + kBlock - 1) / kBlock;
}

static int rng_kind(const tqf_plan* plan) {
  switch (plan->rng.type) {
    case TQF_RNG_PHILOX: return RNGK_PHILOX;
    case TQF_RNG_SOBOL: return RNGK_SOBOL;
    default: return RNGK_DRAWS;
  }
}

// record_dev <- table, unless it already holds exactly that (uploaded on the same stream).
static cudaError_t upload_record(tqf_plan* plan, const int* table, size_t n, cudaStream_t stream) {
  if (plan->record_cache && plan->record_stream == stream && plan->record_cache->size() == n &&
      std::memcmp(plan->record_cache->data(), table, n * sizeof(int)) == 0)
    return cudaSuccess;
  const cudaError_t e =
      cudaMemcpyAsync(plan->record_dev, table, n * sizeof(int), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  if (!plan->record_cache) plan->record_cache = new (std::nothrow) std::vector<int>();
  if (plan->record_cache) {
    plan->record_cache->assign(table, table + n);
    plan->record_stream = stream;
  }
  return cudaSuccess;
}

// Kernel-side descriptor of the next exchange of this plan (advances the epoch).
static PeerK next_peer_exchange(tqf_plan* plan) {
  PeerK pk;
  std::memset(&pk, 0, sizeof(pk));
  if (plan->peer.world > 1) {
    pk.peer_rank = plan->peer.rank;
    pk.peer_world = plan->peer.world;
    pk.peer_epoch = ++plan->peer.epoch;
    for (int r = 0; r < plan->peer.world; ++r) pk.peer_bufs[r] = plan->peer.bufs[r];
  }
  return pk;
}

template <typename Real>
static int dispatch(const tqf_plan* plan, int mode, int grid, size_t smem, const KParams<Real>& P,
                    cudaStream_t stream, int* grid_out) {
  const int rk = rng_kind(plan);
  const bool anti = plan->rng.antithetic != 0;
  switch (plan->model.kind) {
    case TQF_MODEL_AFFINE_1F:
      return launch_path_kernel<AffineModel1F<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_AFFINE_1F_TANGENT:
      return launch_path_kernel<TangentAffine1FModel<Real>>(rk, anti, mode, grid, smem, P, stream,
                                                            grid_out);
    case TQF_MODEL_MILSTEIN_1F:
      return launch_path_kernel<MilsteinAffine1FModel<Real>>(rk, anti, mode, grid, smem, P, stream,
                                                             grid_out);
    case TQF_MODEL_GBM_1F:
      return launch_path_kernel<GbmModel1F<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_LINEAR_1F:
      return launch_path_kernel<LinearModel1F<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_HESTON_EULER:
      return launch_path_kernel<HestonEulerModel<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_HESTON_TANGENT:
      return launch_path_kernel<TangentHestonModel<Real>>(rk, anti, mode, grid, smem, P, stream,
                                                          grid_out);
    case TQF_MODEL_HW1F:
      return launch_path_kernel<HullWhite1FModel<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_HESTON_QE:
      return launch_path_kernel<HestonQeModel<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_AFFINE_ND:
      if (plan->info.dim == 2)
        return launch_path_kernel<AffineModel2D<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      if (plan->info.dim == 3)
        return launch_path_kernel<AffineModel3D<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      return launch_path_kernel<AffineModel4D<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    case TQF_MODEL_HJM: {
      const int f = plan->info.dim - 1, nfs = plan->info.nf;
      if (f == 1 && nfs == 1)
        return launch_path_kernel<HjmModel11<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      if (f == 1)
        return launch_path_kernel<HjmModel12<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      if (f == 2 && nfs == 2)
        return launch_path_kernel<HjmModel22<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      if (f == 2)
        return launch_path_kernel<HjmModel26<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
      return launch_path_kernel<HjmModel33<Real>>(rk, anti, mode, grid, smem, P, stream, grid_out);
    }
    default:
      set_error("model kind not supported by the generic path kernel");
      return TQF_ERR_UNSUPPORTED;
  }
}

template <typename Real>
static int run_price(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                     const tqf_payoff_desc* payoffs, int num_payoffs, double* sums_dev,
                     cudaStream_t stream) {
  KParams<Real> P;
  fill_common(plan, path_offset, path_count, &P);
  P.num_payoffs = num_payoffs;
  int monitor = -1;
  const int S = plan->model.num_steps;
  std::vector<int> flags(static_cast<size_t>(S) + 1, -1);
  std::vector<SwaptionK> swaptions;
  for (int q = 0; q < num_payoffs; ++q) {
    const tqf_payoff_desc& d = payoffs[q];
    TQF_REQUIRE(d.kind >= TQF_PAYOFF_CALL && d.kind <= TQF_PAYOFF_PUT_TANGENT,
                "unknown payoff kind");
    const int step = d.expiry_step > 0 ? d.expiry_step : S;
    TQF_REQUIRE(step <= S, "payoff expiry_step exceeds the number of steps");
    flags[step] = 1;
    if (d.kind == TQF_PAYOFF_HW_SWAPTION) {
      TQF_REQUIRE(plan->model.kind == TQF_MODEL_HW1F || plan->model.kind == TQF_MODEL_HJM,
                  "TQF_PAYOFF_HW_SWAPTION needs the TQF_MODEL_HW1F or TQF_MODEL_HJM model");
      const int nf = d.num_factors > 0 ? d.num_factors : 1;
      TQF_REQUIRE(nf == (plan->model.kind == TQF_MODEL_HJM ? plan->info.dim - 1 : 1),
                  "payoff num_factors does not match the model");
      TQF_REQUIRE(d.num_payments >= 1 && d.num_payments * nf <= TQF_MAX_SWAPTION_PAYMENTS,
                  "num_payments out of range");
      if (swaptions.empty()) swaptions.resize(TQF_MAX_PAYOFFS);
      SwaptionK& sw = swaptions[q];
      std::memset(&sw, 0, sizeof(sw));
      sw.num_payments = d.num_payments;
      sw.is_payer = d.is_payer;
      sw.num_factors = nf;
      for (int j = 0; j < d.num_payments * nf; ++j) sw.g[j] = d.pay_g[j];
      for (int j = 0; j < d.num_payments; ++j) {
        sw.k[j] = d.pay_k[j];
        sw.coef[j] = d.pay_coef[j];
      }
      P.pay[q] = PayoffK{d.kind, 0, 0, step, 0.0, 0.0, d.scale, 0, 0};
      continue;
    }
    if (plan->model.kind == TQF_MODEL_MVGBM) {
      TQF_REQUIRE(d.component >= -1 && d.component < plan->info.dim,
                  "payoff component out of range (-1 = basket mean)");
      TQF_REQUIRE(d.kind == TQF_PAYOFF_CALL || d.kind == TQF_PAYOFF_PUT ||
                      d.kind == TQF_PAYOFF_IDENTITY,
                  "MVGBM supports call / put / identity payoffs");
    } else {
      TQF_REQUIRE(d.component >= 0 && d.component < plan->info.dim,
                  "payoff component out of range");
    }
    const bool is_tangent = d.kind == TQF_PAYOFF_CALL_TANGENT || d.kind == TQF_PAYOFF_PUT_TANGENT;
    TQF_REQUIRE(!is_tangent || (plan->model.kind != TQF_MODEL_MVGBM && d.tangent_component >= 0 &&
                                d.tangent_component < plan->info.dim),
                "tangent_component out of range");
    const bool is_barrier = d.kind >= TQF_PAYOFF_UP_OUT_CALL && d.kind <= TQF_PAYOFF_DOWN_OUT_CALL;
    const int bridge = is_barrier && d.brownian_bridge ? 1 : 0;
    P.pay[q] = PayoffK{d.kind, d.component, d.transform, step, d.strike, d.barrier, d.scale,
                       is_tangent ? d.tangent_component : 0, bridge};
    if (is_barrier) {
      TQF_REQUIRE(monitor < 0 || monitor == d.component,
                  "all barrier payoffs of one call must watch the same state component");
      monitor = d.component;
      const bool up = d.kind == TQF_PAYOFF_UP_OUT_CALL || d.kind == TQF_PAYOFF_UP_OUT_PUT;
      P.need_extrema |= up ? 1 : 2;
      if (bridge) {
        TQF_REQUIRE(d.component == 0, "the Brownian-bridge correction monitors state component 0");
        TQF_REQUIRE(plan->model.kind == TQF_MODEL_AFFINE_1F || plan->model.kind == TQF_MODEL_LINEAR_1F ||
                        plan->model.kind == TQF_MODEL_HESTON_EULER,
                    "the Brownian-bridge correction is implemented for the 1-d affine / additive "
                    "models and the Heston Euler scheme (state component 0)");
        TQF_REQUIRE(d.barrier > 0.0 || d.transform != TQF_TRANSFORM_EXP,
                    "a barrier on exp(state) must be positive");
        // the bridge lives in state space: a barrier on exp(X) is log(barrier) on X
        const double level = d.transform == TQF_TRANSFORM_EXP ? std::log(d.barrier) : d.barrier;
        double* slot = up ? &P.bridge_up : &P.bridge_dn;
        TQF_REQUIRE(!(P.bridge & (up ? 1 : 2)) || *slot == level,
                    "all bridged barrier payoffs of one direction must share one barrier level");
        *slot = level;
        P.bridge |= up ? 1 : 2;
      }
    }
  }
  P.monitor = monitor < 0 ? 0 : monitor;
  TQF_CUDA_OK(upload_record(plan, flags.data(), flags.size(), stream));
  P.record_slot = plan->record_dev;
  if (!swaptions.empty()) {
    TQF_CUDA_OK(cudaMemcpyAsync(plan->swaptions_dev, swaptions.data(),
                                swaptions.size() * sizeof(SwaptionK), cudaMemcpyHostToDevice,
                                stream));
    // the staging vector dies at return: pageable copies are complete on return
  }
  P.swaptions = plan->swaptions_dev;
  P.partials = plan->partials_dev;
  int grid = 1;
  if (plan->model.kind == TQF_MODEL_MVGBM) {
    TQF_REQUIRE(!plan->rng.antithetic && plan->rng.type != TQF_RNG_DRAWS,
                "MVGBM supports the Philox and Sobol generators without antithetic pairing");
    MvLaunch a;
    std::memset(&a, 0, sizeof(a));
    a.dtype = plan->model.dtype;
    a.dim = plan->info.dim;
    a.num_steps = S;
    a.num_steps_total = plan->model.num_steps_total;
    a.rngk = rng_kind(plan);
    a.mode = MODE_PRICE;
    a.max_grid = plan->max_grid;
    a.coef_dev = plan->coef_dev;
    a.x0 = plan->x0;
    a.mu = plan->mu;
    a.sigma = plan->sigma;
    a.chol = plan->chol;
    a.key = P.key;
    a.ctr = P.ctr;
    a.sobol_v = plan->sobol_dev;
    a.logtab = plan->logtab_dev;
    a.ndtab = plan->ndtab_dev;
    a.lsplit_dev = plan->lsplit_dev;
    a.first_index = P.first_index;
    a.path_offset = path_offset;
    a.path_count = path_count;
    a.num_payoffs = num_payoffs;
    a.pay = P.pay;
    a.partials = plan->partials_dev;
    a.record_dev = plan->record_dev;
    a.exact_log = plan->model.reserved;
    a.sobol_clamp = plan->sobol_clamp;
    int rc = launch_mvgbm(a, stream, &grid);
    if (rc != TQF_OK) return rc;
    reduce_partials_kernel<<<1, 1024, 0, stream>>>(plan->partials_dev, grid, num_payoffs, sums_dev,
                                                  next_peer_exchange(plan));
    TQF_CUDA_OK(cudaGetLastError());
    return TQF_OK;
  }
  const int rk = rng_kind(plan);
  bool in_smem = true;
  size_t smem = path_kernel_smem<Real>(plan->info.ncoef, P.num_steps, rk, MODE_PRICE, true);
  if (smem > 96 * 1024) {
    in_smem = false;
    smem = path_kernel_smem<Real>(plan->info.ncoef, P.num_steps, rk, MODE_PRICE, false);
  }
  P.tables_in_smem = in_smem ? 1 : 0;
  const int mode = P.bridge ? MODE_PRICE_BRIDGE : (P.need_extrema ? MODE_PRICE_EXTREMA : MODE_PRICE);
  int rc = dispatch<Real>(plan, mode, plan->max_grid, smem, P, stream, &grid);
  if (rc != TQF_OK) return rc;
  reduce_partials_kernel<<<1, 1024, 0, stream>>>(plan->partials_dev, grid, num_payoffs, sums_dev,
                                                  next_peer_exchange(plan));
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

// sums[m] = sum_blocks partials[block][m]: one warp per column, lanes stride
// over the CTAs, fixed order -> reproducible.
__global__ void colsum_reduce_kernel(const double* __restrict__ partials, int num_blocks, int M,
                                     double* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  double v = 0.0;
  for (int bk = lane; bk < num_blocks; bk += 32) v += partials[static_cast<size_t>(bk) * M + m];
  v = warp_sum(v);
  if (lane == 0) sums[m] = v;
}

template <typename Real>
static int run_paths(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                     const int32_t* record_slot, void* out_dev, int64_t stride_path,
                     int64_t stride_time, int64_t stride_dim, int transform,
                     cudaStream_t stream, int num_slots = 0, double* column_sums_dev = nullptr) {
  KParams<Real> P;
  fill_common(plan, path_offset, path_count, &P);
  if (column_sums_dev) {
    TQF_REQUIRE(plan->model.kind != TQF_MODEL_MVGBM,
                "column sums are not available for the multi-asset kernel");
    const int cols = num_slots * plan->info.dim;
    TQF_REQUIRE(num_slots >= 1 && cols <= 2048, "column sums: 1 <= num_slots * dim <= 2048");
    const size_t need = static_cast<size_t>(plan->max_grid) * cols;
    if (need > plan->colsum_doubles) {
      void* old = plan->colsum_dev;
      dev_release(&old, 1);
      plan->colsum_dev = nullptr;
      plan->colsum_doubles = 0;
      void* fresh = nullptr;
      const int rc = dev_alloc(&fresh, need * sizeof(double));   // block cache: no cudaMalloc per plan
      if (rc != TQF_OK) return rc;
      plan->colsum_dev = static_cast<double*>(fresh);
      plan->colsum_doubles = need;
    }
    P.colsum_partials = plan->colsum_dev;
    P.colsum_cols = cols;
  }
  const size_t nrec = static_cast<size_t>(plan->model.num_steps) + 1;
  TQF_CUDA_OK(upload_record(plan, record_slot, nrec, stream));
  P.record_slot = plan->record_dev;
  P.out = static_cast<Real*>(out_dev);
  P.stride_path = stride_path;
  P.stride_time = stride_time;
  P.stride_dim = stride_dim;
  P.store_exp = transform == TQF_TRANSFORM_EXP ? 1 : 0;
  if (plan->model.kind == TQF_MODEL_MVGBM) {
    TQF_REQUIRE(!plan->rng.antithetic && plan->rng.type != TQF_RNG_DRAWS,
                "MVGBM supports the Philox and Sobol generators without antithetic pairing");
    MvLaunch a;
    std::memset(&a, 0, sizeof(a));
    a.dtype = plan->model.dtype;
    a.dim = plan->info.dim;
    a.num_steps = plan->model.num_steps;
    a.num_steps_total = plan->model.num_steps_total;
    a.rngk = rng_kind(plan);
    a.mode = MODE_PATHS;
    a.max_grid = plan->max_grid;
    a.coef_dev = plan->coef_dev;
    a.x0 = plan->x0;
    a.mu = plan->mu;
    a.sigma = plan->sigma;
    a.chol = plan->chol;
    a.key = P.key;
    a.ctr = P.ctr;
    a.sobol_v = plan->sobol_dev;
    a.logtab = plan->logtab_dev;
    a.ndtab = plan->ndtab_dev;
    a.lsplit_dev = plan->lsplit_dev;
    a.first_index = P.first_index;
    a.path_offset = path_offset;
    a.path_count = path_count;
    a.record_dev = plan->record_dev;
    a.out = out_dev;
    a.stride_path = stride_path;
    a.stride_time = stride_time;
    a.stride_dim = stride_dim;
    a.store_exp = transform == TQF_TRANSFORM_EXP ? 1 : 0;
    a.exact_log = plan->model.reserved;
    a.sobol_clamp = plan->sobol_clamp;
    int g = 1;
    return launch_mvgbm(a, stream, &g);
  }
  P.anti_half = path_count;  // rows of the antithetic partners follow the shard's own rows
  int grid = 1;
  const int rk = rng_kind(plan);
  bool in_smem = true;
  size_t smem = path_kernel_smem<Real>(plan->info.ncoef, P.num_steps, rk, MODE_PATHS, true);
  if (smem > 96 * 1024) {
    in_smem = false;
    smem = path_kernel_smem<Real>(plan->info.ncoef, P.num_steps, rk, MODE_PATHS, false);
  }
  P.tables_in_smem = in_smem ? 1 : 0;
  int rc = dispatch<Real>(plan, MODE_PATHS, plan->max_grid, smem, P, stream, &grid);
  if (rc == TQF_OK && column_sums_dev) {
    colsum_reduce_kernel<<<(P.colsum_cols + 3) / 4, 128, 0, stream>>>(
        plan->colsum_dev, grid, P.colsum_cols, column_sums_dev);
    TQF_CUDA_OK(cudaGetLastError());
  }
  return rc;
}

extern "C" {

int tqf_plan_create(const tqf_model_desc* model, const tqf_rng_desc* rng,
                    uint64_t num_paths_total, tqf_plan** out_plan) {
  TQF_NVTX("tqf_plan_create");
  TQF_REQUIRE(model && rng && out_plan, "null argument");
  *out_plan = nullptr;
  ModelInfo info;
  if (!model_info(model->kind, model->dim, model->num_factors, &info)) {
    set_error("unknown model kind");
    return TQF_ERR_UNSUPPORTED;
  }
  TQF_REQUIRE(model->dtype == TQF_F32 || model->dtype == TQF_F64, "bad dtype");
  TQF_REQUIRE(model->dim == info.dim && model->num_factors == info.nf &&
                  model->num_coef == info.ncoef,
              "dim / num_factors / num_coef do not match the model kind");
  TQF_REQUIRE(model->num_steps >= 0 && model->num_steps <= model->num_steps_total,
              "num_steps must be in [0, num_steps_total]");
  TQF_REQUIRE(model->num_steps == 0 || model->coef, "null coefficient table");
  TQF_REQUIRE(model->x0, "null initial state");
  TQF_REQUIRE(model->x0_paths_dev == nullptr || model->kind != TQF_MODEL_MVGBM,
              "per-path initial states are not supported by the multi-asset kernels");
  TQF_REQUIRE(rng->type == TQF_RNG_PHILOX || rng->type == TQF_RNG_SOBOL ||
                  rng->type == TQF_RNG_DRAWS,
              "unknown rng type");
  if (rng->type == TQF_RNG_PHILOX)
    TQF_REQUIRE((rng->counter[1] >> 31) == 0,
                "Philox counter: the low 64 bits must start below 2^63 (TensorFlow's seed "
                "derivations start them at 0)");
  if (rng->antithetic) {
    TQF_REQUIRE(rng->type == TQF_RNG_PHILOX, "antithetic sampling needs the Philox generator");
    TQF_REQUIRE(num_paths_total % 2 == 0,
                "First dimension of `sample_shape` should be even for PSEUDO_ANTITHETIC random type");
  }
  const uint64_t dims = static_cast<uint64_t>(model->num_steps_total) * info.nf;
  if (rng->type == TQF_RNG_SOBOL) {
    TQF_REQUIRE(rng->direction_numbers, "null direction numbers");
    TQF_REQUIRE(dims <= 21201, "Sobol dimension (steps * factors) exceeds 21201");
    TQF_REQUIRE(rng->skip + rng->unit_offset + num_paths_total < 2147483647ull,
                "skip + num_samples too large");
  }
  if (rng->type == TQF_RNG_DRAWS) TQF_REQUIRE(rng->draws_dev, "null normal_draws");
  TQF_REQUIRE(rng->type != TQF_RNG_SOBOL || rng->unit_stride <= 1,
              "Sobol draws need unit_stride == 1");
  TQF_REQUIRE(model->kind != TQF_MODEL_MVGBM || (rng->unit_stride <= 1 && rng->unit_offset == 0),
              "MVGBM does not support batched draw units");

  int ndev = 0;
  TQF_CUDA_OK(cudaGetDeviceCount(&ndev));
  if (ndev == 0) {
    set_error("no CUDA device: libtqf has no CPU fallback");
    return TQF_ERR_CUDA;
  }
  tqf_plan* plan = new (std::nothrow) tqf_plan();
  TQF_REQUIRE(plan, "out of memory");
  std::memset(plan, 0, sizeof(*plan));
  plan->model = *model;
  plan->rng = *rng;
  plan->info = info;
  plan->num_paths_total = num_paths_total;
  for (int j = 0; j < info.dim; ++j) plan->x0[j] = model->x0[j];
  if (model->kind == TQF_MODEL_MVGBM) {
    if (!model->matrix || !model->vector) {
      delete plan;
      set_error("MVGBM needs the Cholesky factor (matrix) and means / volatilities (vector)");
      return TQF_ERR_INVALID_ARGUMENT;
    }
    for (int i = 0; i < info.dim; ++i) {
      plan->mu[i] = model->vector[i];
      plan->sigma[i] = model->vector[info.dim + i];
      for (int j = 0; j < info.dim; ++j)
        plan->chol[i * info.dim + j] = model->matrix[static_cast<size_t>(i) * info.dim + j];
    }
  }
  int rc = TQF_OK;
  cudaError_t e = cudaGetDevice(&plan->device);
  int sms = kSMs;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device);
  if (e != cudaSuccess) rc = cuda_fail(e, "cudaGetDevice");
  plan->max_grid = sms * grid_per_sm();
  if (rc == TQF_OK) {
    rc = model->dtype == TQF_F64 ? upload_coef<double>(model, info.ncoef, &plan->coef_dev)
                                 : upload_coef<float>(model, info.ncoef, &plan->coef_dev);
  }
  if (rc == TQF_OK && rng->type == TQF_RNG_SOBOL)
    rc = upload_sobol_table(rng->direction_numbers, static_cast<int>(dims), &plan->sobol_dev, 0);
  if (rc == TQF_OK) rc = device_logtab(&plan->logtab_dev);
  if (rc == TQF_OK && model->kind == TQF_MODEL_MVGBM) rc = device_ndtri_f32_tab(&plan->ndtab_dev);
  if (rc == TQF_OK && model->kind == TQF_MODEL_MVGBM && info.dim > 8)
    rc = mvgbm_upload_split(plan->chol, plan->mu, plan->sigma, info.dim, model->dtype,
                            &plan->lsplit_dev);
  if (rc == TQF_OK) {
    void* p = nullptr;
    rc = dev_alloc(&p, static_cast<size_t>(plan->max_grid) * TQF_MAX_PAYOFFS * 4 * sizeof(double));
    plan->partials_dev = static_cast<double*>(p);
    if (rc == TQF_OK) {
      p = nullptr;
      rc = dev_alloc(&p, (static_cast<size_t>(model->num_steps) + 1) * sizeof(int));
      plan->record_dev = static_cast<int*>(p);
    }
    if (rc == TQF_OK) {
      p = nullptr;
      rc = dev_alloc(&p, TQF_MAX_PAYOFFS * sizeof(SwaptionK));
      plan->swaptions_dev = static_cast<SwaptionK*>(p);
    }
  }
  // the descriptors' host pointers are not kept
  plan->model.coef = nullptr;
  plan->model.x0 = nullptr;
  plan->model.matrix = nullptr;
  plan->model.vector = nullptr;
  plan->rng.direction_numbers = nullptr;
  if (rc != TQF_OK) {
    tqf_plan_destroy(plan);
    return rc;
  }
  *out_plan = plan;
  return TQF_OK;
}

int tqf_plan_destroy(tqf_plan* plan) {
  if (!plan) return TQF_OK;
  void* blocks[7] = {plan->coef_dev,     plan->sobol_dev,  plan->lsplit_dev,   plan->colsum_dev,
                     plan->partials_dev, plan->record_dev, plan->swaptions_dev};
  // (the blocks belong to the plan's device, whichever device is current in this thread)
  int current = -1;
  cudaGetDevice(&current);
  if (current != plan->device) cudaSetDevice(plan->device);
  dev_release(blocks, 7);
  if (current >= 0 && current != plan->device) cudaSetDevice(current);
  delete plan->record_cache;
  delete plan;
  return TQF_OK;
}

int tqf_plan_price(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                   const tqf_payoff_desc* payoffs, int num_payoffs, double* sums_dev,
                   void* stream) {
  TQF_NVTX("tqf_plan_price");
  TQF_REQUIRE(plan && sums_dev, "null argument");
  TQF_REQUIRE(num_payoffs >= 1 && num_payoffs <= TQF_MAX_PAYOFFS && payoffs,
              "num_payoffs must be in [1, TQF_MAX_PAYOFFS]");
  const uint64_t units = plan->rng.antithetic ? plan->num_paths_total / 2 : plan->num_paths_total;
  TQF_REQUIRE(path_offset + path_count <= units, "shard exceeds the number of paths");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return plan->model.dtype == TQF_F64
             ? run_price<double>(plan, path_offset, path_count, payoffs, num_payoffs, sums_dev, s)
             : run_price<float>(plan, path_offset, path_count, payoffs, num_payoffs, sums_dev, s);
}

int tqf_plan_price_host(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                        const tqf_payoff_desc* payoffs, int num_payoffs, double* sums_dev,
                        double* sums_host, void* stream) {
  TQF_NVTX("tqf_plan_price_host");
  TQF_REQUIRE(sums_host, "null host buffer");
  const int rc = tqf_plan_price(plan, path_offset, path_count, payoffs, num_payoffs, sums_dev, stream);
  if (rc != TQF_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TQF_CUDA_OK(cudaMemcpyAsync(sums_host, sums_dev, sizeof(double) * 4 * num_payoffs,
                              cudaMemcpyDeviceToHost, s));
  TQF_CUDA_OK(cudaStreamSynchronize(s));
  return TQF_OK;
}

int tqf_plan_paths(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                   const int32_t* record_slot, void* out_dev, int64_t stride_path,
                   int64_t stride_time, int64_t stride_dim, int transform, void* stream) {
  TQF_NVTX("tqf_plan_paths");
  TQF_REQUIRE(plan && record_slot, "null argument");
  const uint64_t units = plan->rng.antithetic ? plan->num_paths_total / 2 : plan->num_paths_total;
  TQF_REQUIRE(path_offset + path_count <= units, "shard exceeds the number of paths");
  if (path_count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return plan->model.dtype == TQF_F64
             ? run_paths<double>(plan, path_offset, path_count, record_slot, out_dev, stride_path,
                                 stride_time, stride_dim, transform, s)
             : run_paths<float>(plan, path_offset, path_count, record_slot, out_dev, stride_path,
                                stride_time, stride_dim, transform, s);
}

int tqf_plan_set_peer_exchange(tqf_plan* plan, int rank, int world, void* const* bufs,
                               uint64_t epoch_base) {
  TQF_REQUIRE(plan && bufs, "null argument");
  TQF_REQUIRE(world >= 1 && world <= kLsmMaxPeers && rank >= 0 && rank < world,
              "peer exchange supports up to 8 ranks");
  plan->peer.rank = rank;
  plan->peer.world = world;
  plan->peer.epoch = epoch_base;
  for (int r = 0; r < world; ++r) {
    TQF_REQUIRE(bufs[r] != nullptr, "null peer buffer");
    plan->peer.bufs[r] = static_cast<unsigned char*>(bufs[r]);
  }
  return TQF_OK;
}

int tqf_plan_set_sobol_clamp(tqf_plan* plan, int clamp) {
  TQF_REQUIRE(plan, "null argument");
  plan->sobol_clamp = clamp ? 1 : 0;
  return TQF_OK;
}

int tqf_plan_peer_epoch(const tqf_plan* plan, uint64_t* epoch) {
  TQF_REQUIRE(plan && epoch, "null argument");
  *epoch = plan->peer.epoch;
  return TQF_OK;
}

int tqf_plan_paths_sums(tqf_plan* plan, uint64_t path_offset, uint64_t path_count,
                        const int32_t* record_slot, void* out_dev, int64_t stride_path,
                        int64_t stride_time, int64_t stride_dim, int transform, int num_slots,
                        double* column_sums_dev, void* stream) {
  TQF_NVTX("tqf_plan_paths_sums");
  TQF_REQUIRE(plan && record_slot && column_sums_dev, "null argument");
  const uint64_t units = plan->rng.antithetic ? plan->num_paths_total / 2 : plan->num_paths_total;
  TQF_REQUIRE(path_offset + path_count <= units, "shard exceeds the number of paths");
  TQF_REQUIRE(path_count > 0 && out_dev, "empty shard or null output");
  for (int i = 0; i <= plan->model.num_steps; ++i)
    TQF_REQUIRE(record_slot[i] < num_slots, "record_slot entry exceeds num_slots");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return plan->model.dtype == TQF_F64
             ? run_paths<double>(plan, path_offset, path_count, record_slot, out_dev, stride_path,
                                 stride_time, stride_dim, transform, s, num_slots, column_sums_dev)
             : run_paths<float>(plan, path_offset, path_count, record_slot, out_dev, stride_path,
                                stride_time, stride_dim, transform, s, num_slots, column_sums_dev);
}

}  // extern "C"
