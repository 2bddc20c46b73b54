// Longstaff-Schwartz backward induction as ONE persistent cooperative kernel.
//
// Replaces the per-date loop of models/longstaff_schwartz/lsm.py:296-330
// (`_lsm_loop_body` 403-436: payoff_fn, basis_fn, masked X'X / X'y matmuls,
// tf.linalg.pinv, tf.where updates -- a dozen full passes over [N] tensors per
// exercise date in the reference) for the single-asset American put of config
// C5: time-major contiguous paths, polynomial basis of at most 6 functions, one
// payoff.  tqf_lsm.cu runs the same algorithm with one launch per date; here
//   * one CTA per SM owns a fixed, contiguous set of path tiles for the whole
//     induction; per date it streams the two path columns the date needs
//     and the merged state W = cashflow + values through a 6-stage
//     shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx; a producer warp runs ahead of the 11 consumer warps -- for
//     the read-only columns across date boundaries as well, for W up to the
//     date boundary, where it waits for the consumers' stores), and writes W
//     back with 16-byte stores tagged evict_last so that it stays in the
//     126 MB L2 between dates;
//   * consecutive dates sweep the tiles in opposite directions: the column a
//     date accumulates on is the column the next date updates with, and its
//     most recently read tail is still in L2 when the sweep turns around;
//   * the per-date dependency (beta of date e needs the normal equations of ALL
//     paths) is a grid barrier: every CTA publishes its 27 partial sums, the
//     last CTA to arrive reduces them in a fixed order, exchanges them with the
//     other GPUs of the box over NVLink peer memory when the paths are sharded
//     (tqf_peer.cuh), solves the K x K system once and releases a flag the
//     other CTAs spin on;
//   * the initial cashflow (first sweep) and the final value sum (last sweep)
//     are fused in: the whole LSM part of C5 is one launch instead of 52.
//
// HBM/L2-bound: 32 algorithmic bytes per path and date (two columns, read +
// write of W), of which W is served by L2.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "tqf_lsm_internal.cuh"

namespace tqf {

constexpr int kPConsumers = 352;                  // 11 consumer warps (12 warps per CTA: 168 registers each)
constexpr int kPThreads = kPConsumers + 32;       // + one producer warp
constexpr int kPVecPerThread = 2;                 // 16-byte vectors per thread and tile
constexpr int kPTileVecs = kPConsumers * kPVecPerThread;   // 704 vectors = 11 KB per column
constexpr int kPTileBytes = kPTileVecs * 16;
constexpr int kPStages = 6;                       // 6 x 3 x 11 KB = 198 KB of columns and W in flight
constexpr int kPSlots = 3;                        // update column | accumulation column | W
constexpr size_t kPSmemBytes = static_cast<size_t>(kPStages) * kPSlots * kPTileBytes + 1024;
constexpr long long kPTimeoutCycles = 8000000000ll;   // ~4 s: a lost CTA must not hang the GPU

template <typename Real>
struct PersistArgs {
  const Real* paths;          // time-major: column t at paths + t * stride_time
  int64_t stride_time;
  Real* w;                    // [N]
  uint32_t num_vecs;          // N / (16 / sizeof(Real))
  uint64_t path_offset, num_calib, skip_below;
  const int* ex_times;        // device [T]
  int T;
  const double* means;        // device: mean of exercise slot s at means[s]
  const double* ratio;        // device [T]: row e = df[e + 1] / df[e]
  double strike;
  double rcond;
  int round_to_float;
  double* partials;           // [gridDim.x][27]
  double* sums;               // [27] reduced (and exchanged) sums of the current date
  double* beta;               // [K]
  double* history;            // optional [T - 1][27 + 6]: sums and beta of every date
  double* value_sums;         // [2]: sum of W over the pricing paths, their count
  unsigned long long* ctrl;   // [0] arrivals, [1] release flag, [2] status
  int tune;                   // TQF_LSM_TUNE experiment bits (0 = the shipped configuration)
  PeerK peer;
};

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void p_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void p_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void p_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool p_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a barrier that never completes (a bug, a lost peer) must not hang the box.
__device__ __forceinline__ bool p_mbar_wait(uint32_t bar, uint32_t parity, unsigned sleep_ns = 0) {
  if (p_mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!p_mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);       // producer: do not compete for issue slots
    if (clock64() - t0 > kPTimeoutCycles) return false;
  }
  return true;
}
__device__ __forceinline__ void p_bulk_load(uint32_t dst, const void* src, uint32_t bytes,
                                            uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void p_bar_consumers() {
  asm volatile("bar.sync 1, %0;" :: "n"(kPConsumers) : "memory");
}
__device__ __forceinline__ uint4 p_ld_keep(const uint4* ptr, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void p_st_keep(uint4* ptr, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 p_lds(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <typename Real> struct PVec;
template <> struct PVec<double> {
  static constexpr int N = 2;
  __device__ static __forceinline__ void unpack(const uint4& v, double (&x)[2]) {
    x[0] = __hiloint2double(static_cast<int>(v.y), static_cast<int>(v.x));
    x[1] = __hiloint2double(static_cast<int>(v.w), static_cast<int>(v.z));
  }
  __device__ static __forceinline__ uint4 pack(const double (&x)[2]) {
    return make_uint4(static_cast<uint32_t>(__double2loint(x[0])), static_cast<uint32_t>(__double2hiint(x[0])),
                      static_cast<uint32_t>(__double2loint(x[1])), static_cast<uint32_t>(__double2hiint(x[1])));
  }
};
template <> struct PVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&x)[4]) {
    x[0] = __uint_as_float(v.x);
    x[1] = __uint_as_float(v.y);
    x[2] = __uint_as_float(v.z);
    x[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&x)[4]) {
    return make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]),
                      __float_as_uint(x[3]));
  }
};

// The K x K solve of the last CTA: not inlined, so that its ~190 registers do
// not weigh on the streaming loop.
template <int KT>
__device__ __noinline__ void p_solve(const double* sums, double rcond, int round_to_float,
                                     double* beta) {
  lsm_solve_one<KT>(sums, rcond, round_to_float, beta);
}

// W' = (pv > 0 and pv > cont) ? pv : rw -- two compares chained through the
// predicate and one select (the plain C++ form compiles to two selects per path)
__device__ __forceinline__ double p_exercise(double pv, double cont, double rw) {
  double r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f64 q, %1, 0d0000000000000000;\n\t"
      "setp.gt.and.f64 p, %1, %2, q;\n\tselp.f64 %0, %1, %3, p;\n\t}"
      : "=d"(r) : "d"(pv), "d"(cont), "d"(rw));
  return r;
}
__device__ __forceinline__ float p_exercise(float pv, float cont, float rw) {
  float r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.gt.f32 q, %1, 0f00000000;\n\t"
      "setp.gt.and.f32 p, %1, %2, q;\n\tselp.f32 %0, %1, %3, p;\n\t}"
      : "=f"(r) : "f"(pv), "f"(cont), "f"(rw));
  return r;
}

struct SweepState {
  uint32_t stage, phase;     // position in the shared-memory ring (continues across dates)
  uint32_t wpar;             // per stage: parity of its W barrier (unused on the first date)
  bool failed;
};
template <typename Real, int KT>
struct SweepParams {
  Real beta[KT];
  Real strike, mean_u, ratio_u, mean_a, ratio_a;
  bool descending;
};

// One date of the backward induction over this CTA's tiles (consumer threads).
// MID = true: the common sweep (update + accumulate, every path calibrates),
// straight-line code per tile; MID = false: the first / last sweep and
// num_calibration_samples, with run-time flags.  Leaves the K (K + 1) / 2 + K
// regression sums (or, on the last sweep, {value sum, count}) in `acc`.
template <typename Real, int KT, bool MID>
__device__ __forceinline__ void p_sweep(const PersistArgs<Real>& A, const SweepParams<Real, KT>& sp,
                                        SweepState& st, double (&acc)[KT * (KT + 1) / 2 + KT],
                                        uint32_t ring, uint32_t fullb, uint32_t emptyb,
                                        uint32_t fullwb, uint32_t passdone,
                                        uint32_t tile_lo, uint32_t tile_hi, int tid, int lane,
                                        uint64_t keep, bool first, bool last) {
  constexpr int VN = PVec<Real>::N;
  constexpr int NX = KT * (KT + 1) / 2;
  uint4* wv = reinterpret_cast<uint4*>(A.w);
  const uint32_t num_my = tile_hi - tile_lo;
  const bool calib_all = A.num_calib == ~0ull;
  uint32_t count = 0;                           // sum of phi_0 phi_0: an integer
  double vsum = 0.0, vcnt = 0.0;
  uint32_t tile = sp.descending ? tile_hi - 1 : tile_lo;
  const int step = sp.descending ? -1 : 1;
  for (uint32_t i = 0; i < num_my; ++i, tile += step) {
    if (!p_mbar_wait(fullb + 8 * st.stage, st.phase)) st.failed = true;
    if (MID || !first) {
      if (!p_mbar_wait(fullwb + 8 * st.stage, (st.wpar >> st.stage) & 1u)) st.failed = true;
      st.wpar ^= 1u << st.stage;
    }
    const uint32_t src = ring + st.stage * (kPSlots * kPTileBytes);
    uint4 xu_raw[kPVecPerThread], xa_raw[kPVecPerThread], wreg[kPVecPerThread];
#pragma unroll
    for (int u = 0; u < kPVecPerThread; ++u) {
      xu_raw[u] = p_lds(src + u * kPConsumers * 16);
      if (MID || !last) xa_raw[u] = p_lds(src + kPTileBytes + u * kPConsumers * 16);
      if (MID || !first) wreg[u] = p_lds(src + 2 * kPTileBytes + u * kPConsumers * 16);
    }
    __syncwarp();
    if (lane == 0) p_mbar_arrive(emptyb + 8 * st.stage);     // the stage is free again
    if (++st.stage == kPStages) {
      st.stage = 0;
      st.phase ^= 1u;
    }
    const uint32_t vbase = tile * kPTileVecs + tid;
    Real wn[kPVecPerThread][VN];
    // ---- exercise decision: W' = (pv > 0 and pv > X beta) ? pv : ratio W, which is
    // `ev > relu(X beta) ? ev : ratio W` with ev = relu(pv) (lsm.py:391-399)
#pragma unroll
    for (int u = 0; u < kPVecPerThread; ++u) {
      const uint32_t v = vbase + u * kPConsumers;
      // (lanes beyond the last vector of a partial tile compute on stale shared
      // memory; only their store and their sums are masked)
      Real xu[VN], wo[VN];
      PVec<Real>::unpack(xu_raw[u], xu);
      if (MID || !first) PVec<Real>::unpack(wreg[u], wo);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const Real pv = sp.strike - xu[e];
        if (!MID && first) {
          wn[u][e] = pv > Real(0) ? pv : Real(0);              // the terminal cashflow
        } else {
          const Real c = xu[e] - sp.mean_u;
          Real cont = sp.beta[KT - 1];
#pragma unroll
          for (int k = KT - 2; k >= 0; --k) cont = fma(cont, c, sp.beta[k]);
          wn[u][e] = p_exercise(pv, cont, sp.ratio_u * wo[e]);
        }
      }
      if (v < A.num_vecs) p_st_keep(wv + v, PVec<Real>::pack(wn[u]), keep);
    }
    // ---- normal equations of the next (earlier) date, branch-free: a path that does
    // not take part (out of the money) contributes c = 0, y = 0 and no count
#pragma unroll
    for (int u = 0; u < kPVecPerThread; ++u) {
      const uint32_t v = vbase + u * kPConsumers;
      const bool active = v < A.num_vecs;
      if (MID || !last) {
        Real xa[VN];
        PVec<Real>::unpack(xa_raw[u], xa);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const Real pa = sp.strike - xa[e];
          bool use = active && pa > Real(0);
          if (!MID && !calib_all)
            use = use && (A.path_offset + static_cast<uint64_t>(v) * VN + e) < A.num_calib;
          const Real cr = use ? xa[e] - sp.mean_a : Real(0);
          const double y = use ? static_cast<double>(sp.ratio_a * wn[u][e]) : 0.0;
          count += use ? 1u : 0u;
          double phi[KT];
          Real pw = 1;
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            phi[k] = static_cast<double>(pw);
            pw *= cr;
          }
          int idx = 0;
#pragma unroll
          for (int a = 0; a < KT; ++a)
#pragma unroll
            for (int b = a; b < KT; ++b) {
              if (idx > 0) acc[idx] = fma(phi[a], phi[b], acc[idx]);    // phi_0 = 1: additions
              ++idx;
            }
          acc[NX] += y;
#pragma unroll
          for (int a = 1; a < KT; ++a) acc[NX + a] = fma(phi[a], y, acc[NX + a]);
        }
      } else if (active) {
#pragma unroll
        for (int e = 0; e < VN; ++e)
          if (A.path_offset + static_cast<uint64_t>(v) * VN + e >= A.skip_below) {
            vsum += static_cast<double>(wn[u][e]);
            vcnt += 1.0;
          }
      }
    }
  }
  if (MID || !last) {
    acc[0] = static_cast<double>(count);
  } else {
    acc[0] = vsum;
    acc[1] = vcnt;
  }
  // this warp's W stores of the date are complete: make them visible to the bulk
  // copies (async proxy) that fetch W for the next date, then tell the W producer
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  if (lane == 0) p_mbar_arrive(passdone);
}

template <typename Real, int KT>
__global__ void __launch_bounds__(kPThreads, 1) lsm_persistent_kernel(const PersistArgs<Real> A) {
  constexpr int VN = PVec<Real>::N;                   // paths per 16-byte vector
  constexpr int NA = KT * (KT + 1) / 2 + KT;
  extern __shared__ __align__(128) unsigned char p_smem[];
  // layout: [stages][2 columns][16 KB] | full barriers | empty barriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + static_cast<size_t>(kPStages) * kPSlots * kPTileBytes);
  const uint32_t smem_base = p_smem_u32(p_smem);
  // full0[s]: the two path columns of stage s have landed; fullw0[s]: its W tile has;
  // empty0[s]: every consumer warp has copied the stage into registers; passdone: every
  // consumer warp has stored (and fenced) its last W of the current date
  const uint32_t full0 = p_smem_u32(bars), empty0 = p_smem_u32(bars + kPStages);
  const uint32_t fullw0 = p_smem_u32(bars + 2 * kPStages), passdone = p_smem_u32(bars + 3 * kPStages);
  __shared__ double s_red[kPConsumers / 32][32];
  __shared__ double s_beta[kLsmFastK];
  __shared__ int s_flag;

  const int tid = threadIdx.x;
  const int G = gridDim.x, cta = blockIdx.x;
  const uint32_t ntiles = (A.num_vecs + kPTileVecs - 1) / kPTileVecs;
  const uint32_t tile_lo = static_cast<uint32_t>(static_cast<uint64_t>(ntiles) * cta / G);
  const uint32_t tile_hi = static_cast<uint32_t>(static_cast<uint64_t>(ntiles) * (cta + 1) / G);
  const uint32_t num_my = tile_hi - tile_lo;
  const int T = A.T;

  if (tid == 0) {
    for (int s = 0; s < kPStages; ++s) {
      p_mbar_init(full0 + 8 * s, 1);
      p_mbar_init(fullw0 + 8 * s, 1);
      p_mbar_init(empty0 + 8 * s, kPConsumers / 32);
    }
    p_mbar_init(passdone, kPConsumers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kPConsumers) {
    // ------------------------------------------------------------ producer
    if (tid == kPConsumers) {
      // lane 0: the two path columns of every tile, free-running across dates (the
      // columns are read-only), bounded only by the ring
      uint64_t pol_first, pol_normal;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
      if (A.tune & 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_normal));
      if (A.tune & 8) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_first));
      uint32_t it = 0;
      for (int j = 0; j < T; ++j) {
        const bool last = j == T - 1;
        const Real* colu = A.paths + static_cast<int64_t>(A.ex_times[T - 1 - j]) * A.stride_time;
        const Real* cola = last ? colu
                                : A.paths + static_cast<int64_t>(A.ex_times[T - 2 - j]) * A.stride_time;
        for (uint32_t i = 0; i < num_my; ++i, ++it) {
          const uint32_t tile = ((j & 1) && !(A.tune & 4)) ? tile_hi - 1 - i : tile_lo + i;
          const uint32_t s = it % kPStages, ph = (it / kPStages) & 1u;
          if (!p_mbar_wait(empty0 + 8 * s, ph ^ 1u, 200)) {
            atomicExch(A.ctrl + 2, 2ull);
            return;
          }
          const uint32_t v0 = tile * kPTileVecs;
          const uint32_t nv = min(static_cast<uint32_t>(kPTileVecs), A.num_vecs - v0);
          const uint32_t bytes = nv * 16u;
          const uint32_t dst = smem_base + s * (kPSlots * kPTileBytes);
          p_mbar_expect_tx(full0 + 8 * s, last ? bytes : 2 * bytes);
          // the update column is read for the last time: evict_first; the column that
          // is accumulated on comes back as the update column of the next date
          p_bulk_load(dst, reinterpret_cast<const unsigned char*>(colu) + static_cast<size_t>(v0) * 16,
                      bytes, full0 + 8 * s, pol_first);
          if (!last)
            p_bulk_load(dst + kPTileBytes,
                        reinterpret_cast<const unsigned char*>(cola) + static_cast<size_t>(v0) * 16,
                        bytes, full0 + 8 * s, pol_normal);
        }
      }
    } else if (tid == kPConsumers + 1) {
      // lane 1: the W tiles (none on the first date, whose W is the terminal payoff).  W
      // of date j + 1 is what the consumers stored on date j -- and the sweep turns
      // around, so the first tiles wanted are the last ones written: this lane waits at
      // every date boundary until all consumer warps have stored and fenced
      // (`passdone`); within a date the tiles are distinct and it runs ahead like lane 0.
      uint64_t pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      if (A.tune & 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
      uint32_t it = num_my;                     // ring position of the first tile of date 1
      for (int j = 1; j < T; ++j) {
        if (!p_mbar_wait(passdone, static_cast<uint32_t>(j - 1) & 1u, 200)) {
          atomicExch(A.ctrl + 2, 5ull);
          return;
        }
        for (uint32_t i = 0; i < num_my; ++i, ++it) {
          const uint32_t tile = ((j & 1) && !(A.tune & 4)) ? tile_hi - 1 - i : tile_lo + i;
          const uint32_t s = it % kPStages, ph = (it / kPStages) & 1u;
          if (!p_mbar_wait(empty0 + 8 * s, ph ^ 1u, 200)) {
            atomicExch(A.ctrl + 2, 2ull);
            return;
          }
          const uint32_t v0 = tile * kPTileVecs;
          const uint32_t nv = min(static_cast<uint32_t>(kPTileVecs), A.num_vecs - v0);
          const uint32_t dst = smem_base + s * (kPSlots * kPTileBytes) + 2 * kPTileBytes;
          p_mbar_expect_tx(fullw0 + 8 * s, nv * 16u);
          p_bulk_load(dst, reinterpret_cast<const unsigned char*>(A.w) + static_cast<size_t>(v0) * 16,
                      nv * 16u, fullw0 + 8 * s, pol_keep);
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------- consumers
  uint64_t keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  if (A.tune & 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(keep));
  const int warp = tid >> 5, lane = tid & 31;
  SweepState st;
  st.stage = 0;
  st.phase = 0;
  st.wpar = 0;
  st.failed = false;
  // shared-memory addresses computed once (opaque to the compiler, which would
  // otherwise rematerialise the shared-window arithmetic in every iteration)
  uint32_t ring = smem_base + tid * 16, fullb = full0, emptyb = empty0, fullwb = fullw0;
  asm volatile("" : "+r"(ring), "+r"(fullb), "+r"(emptyb), "+r"(fullwb));
  const bool calib_all = A.num_calib == ~0ull;

  for (int j = 0; j < T; ++j) {
    const bool first = j == 0, last = j == T - 1;
    // update of exercise index e_u = T - j on column slot e_u - 1; accumulation for
    // e_a = T - j - 1 on column slot e_a - 1 (lsm.py:403-436)
    SweepParams<Real, KT> sp;
#pragma unroll
    for (int k = 0; k < KT; ++k) sp.beta[k] = first ? Real(0) : static_cast<Real>(s_beta[k]);
    sp.strike = static_cast<Real>(A.strike);
    sp.mean_u = first ? Real(0) : static_cast<Real>(A.means[T - 1 - j]);
    sp.ratio_u = first ? Real(1) : static_cast<Real>(A.ratio[T - j]);
    sp.mean_a = last ? Real(0) : static_cast<Real>(A.means[T - 2 - j]);
    sp.ratio_a = last ? Real(1) : static_cast<Real>(A.ratio[T - 1 - j]);
    sp.descending = (j & 1) != 0 && !(A.tune & 4);
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    if (!first && !last && calib_all)
      p_sweep<Real, KT, true>(A, sp, st, acc, ring, fullb, emptyb, fullwb, passdone, tile_lo, tile_hi,
                              tid, lane, keep, false, false);
    else
      p_sweep<Real, KT, false>(A, sp, st, acc, ring, fullb, emptyb, fullwb, passdone, tile_lo, tile_hi,
                               tid, lane, keep, first, last);

    // ---- CTA partial sums -> global row (packed 6 x 6 layout), fixed order
    const int M = last ? 2 : kLsmFastNS;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) s_red[warp][i] = v;
    }
    p_bar_consumers();
    if (tid < M) {
      int srcslot = tid;
      if (!last) {
        srcslot = -1;
        if (tid < kLsmFastK * (kLsmFastK + 1) / 2) {
          int a = 0, rem = tid;
          while (rem >= kLsmFastK - a) {
            rem -= kLsmFastK - a;
            ++a;
          }
          const int b = a + rem;
          if (a < KT && b < KT) srcslot = a * KT - a * (a - 1) / 2 + (b - a);
        } else {
          const int a = tid - kLsmFastK * (kLsmFastK + 1) / 2;
          if (a < KT) srcslot = KT * (KT + 1) / 2 + a;
        }
      }
      double v = 0.0;
      if (srcslot >= 0)
        for (int wi = 0; wi < kPConsumers / 32; ++wi) v += s_red[wi][srcslot];
      A.partials[static_cast<size_t>(cta) * kLsmFastNS + tid] = v;
    }
    __threadfence();
    p_bar_consumers();
    // ---- grid barrier: the last CTA to arrive reduces, exchanges, solves, releases
    if (tid == 0) {
      const unsigned long long t = atomicAdd(A.ctrl, 1ull);
      s_flag = (t == static_cast<unsigned long long>(j + 1) * G - 1) ? 1 : 0;
    }
    p_bar_consumers();
    if (s_flag) {
      __threadfence();
      // thread t sums column m = t % 32 over the rows t / 32, t / 32 + 16, ...; the 16 row
      // groups are then combined in a fixed order -> reproducible sums
      double a = 0.0;
      if (lane < M) {
#pragma unroll 5
        for (int r = warp; r < G; r += kPConsumers / 32)
          a += __ldcg(A.partials + static_cast<size_t>(r) * kLsmFastNS + lane);
      }
      p_bar_consumers();                   // everybody is done reading s_red (this CTA's partials)
      s_red[warp][lane] = a;
      p_bar_consumers();
      if (warp == 0 && lane < M) {
        double v = 0.0;
#pragma unroll
        for (int wi = 0; wi < kPConsumers / 32; ++wi) v += s_red[wi][lane];
        A.sums[lane] = v;
      }
      __threadfence();
      p_bar_consumers();
      if (A.peer.peer_world > 1) {
        PeerK pk = A.peer;
        pk.peer_epoch = A.peer.peer_epoch + static_cast<unsigned long long>(j) + 1ull;
        if (!peer_all_reduce(pk, A.sums, M, kPConsumers, 1) && tid == 0) atomicExch(A.ctrl + 2, 3ull);
      }
      if (tid == 0) {
        if (!last) {
          p_solve<KT>(A.sums, A.rcond, A.round_to_float, A.beta);
          if (A.history != nullptr) {
            double* hrow = A.history + static_cast<size_t>(j) * (kLsmFastNS + kLsmFastK);
            for (int m = 0; m < kLsmFastNS; ++m) hrow[m] = A.sums[m];
            for (int k = 0; k < kLsmFastK; ++k) hrow[kLsmFastNS + k] = k < KT ? A.beta[k] : 0.0;
          }
        } else {
          A.value_sums[0] = A.sums[0];
          A.value_sums[1] = A.sums[1];
        }
        __threadfence();
        const unsigned long long rel = static_cast<unsigned long long>(j) + 1ull;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(A.ctrl + 1), "l"(rel) : "memory");
      }
    }
    if (last) break;
    if (tid == 0) {
      const unsigned long long want = static_cast<unsigned long long>(j) + 1ull;
      unsigned long long seen = 0;
      const long long t0 = clock64();
      while (true) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(A.ctrl + 1) : "memory");
        if (seen >= want) break;
        if (clock64() - t0 > kPTimeoutCycles) {
          atomicExch(A.ctrl + 2, 1ull);
          break;
        }
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) s_beta[k] = __ldcg(A.beta + k);
    }
    p_bar_consumers();
  }
  if (st.failed && tid == 0) atomicExch(A.ctrl + 2, 4ull);
}

bool lsm_persistent_ok(const tqf_lsm* h) {
  const tqf_lsm_desc& d = h->desc;
  const size_t esz = d.dtype == TQF_F64 ? 8 : 4;
  return lsm_vec_ok(h) && d.batch == 1 && (d.num_paths % (16 / esz)) == 0 &&
         (reinterpret_cast<uintptr_t>(d.paths_dev) % 16) == 0 &&
         (reinterpret_cast<uintptr_t>(h->w_dev) % 16) == 0 &&
         (static_cast<uint64_t>(d.stride_time) * esz) % 16 == 0 && h->ctrl_dev != nullptr;
}

template <typename Real, int KT>
static int launch_persistent(const PersistArgs<Real>& A, int sms, cudaStream_t stream) {
  auto kern = lsm_persistent_kernel<Real, KT>;
  TQF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(kPSmemBytes)));
  int per_sm = 0;
  TQF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPThreads, kPSmemBytes));
  if (per_sm < 1) {
    set_error("the persistent LSM kernel does not fit on this device");
    return TQF_ERR_UNSUPPORTED;
  }
  const uint32_t ntiles = (A.num_vecs + kPTileVecs - 1) / kPTileVecs;
  int grid = static_cast<int>(ntiles < static_cast<uint32_t>(sms) ? ntiles : sms);
  if (grid < 1) grid = 1;
  void* params[] = {const_cast<PersistArgs<Real>*>(&A)};
  // cooperative launch: all CTAs are guaranteed to be co-resident (the grid barrier
  // cannot deadlock); the launch fails instead when they are not
  TQF_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(kPThreads),
                                          params, kPSmemBytes, stream));
  return TQF_OK;
}

template <typename Real>
static int run_persistent_t(tqf_lsm* h, int num_times, const double* means_dev,
                            const double* ratio_dev, double rcond, uint64_t skip_below,
                            double* value_sums_dev, double* beta_dev, double* history_dev,
                            cudaStream_t stream) {
  const tqf_lsm_desc& d = h->desc;
  PersistArgs<Real> A;
  std::memset(&A, 0, sizeof(A));
  A.paths = static_cast<const Real*>(d.paths_dev);
  A.stride_time = d.stride_time;
  A.w = static_cast<Real*>(h->w_dev);
  A.num_vecs = static_cast<uint32_t>(d.num_paths / (16 / sizeof(Real)));
  A.path_offset = d.path_offset;
  A.num_calib = d.num_calibration_samples == 0 ? ~0ull : d.num_calibration_samples;
  A.skip_below = skip_below;
  A.ex_times = h->times_dev;
  A.T = num_times;
  A.means = means_dev;
  A.ratio = ratio_dev;
  A.rcond = rcond;
  A.round_to_float = d.dtype == TQF_F32 ? 1 : 0;
  A.partials = h->partials_dev;
  A.sums = h->partials_dev + static_cast<size_t>(kSMs + 8) * kLsmFastNS;
  A.beta = beta_dev;
  A.history = history_dev;
  A.value_sums = value_sums_dev;
  A.ctrl = h->ctrl_dev;
  if (const char* t = std::getenv("TQF_LSM_TUNE")) A.tune = std::atoi(t);
  A.peer.peer_rank = h->peer_rank;
  A.peer.peer_world = h->peer_world;
  A.peer.peer_epoch = h->peer_epoch;
  for (int r = 0; r < h->peer_world; ++r) A.peer.peer_bufs[r] = h->peer_bufs[r];
  A.strike = h->strike0;
  int dev = 0, sms = kSMs;
  TQF_CUDA_OK(cudaGetDevice(&dev));
  TQF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms > kSMs + 8) sms = kSMs + 8;
  switch (h->K) {
    case 1: return launch_persistent<Real, 1>(A, sms, stream);
    case 2: return launch_persistent<Real, 2>(A, sms, stream);
    case 3: return launch_persistent<Real, 3>(A, sms, stream);
    case 4: return launch_persistent<Real, 4>(A, sms, stream);
    case 5: return launch_persistent<Real, 5>(A, sms, stream);
    default: return launch_persistent<Real, 6>(A, sms, stream);
  }
}

int lsm_run_persistent(tqf_lsm* h, const int32_t* exercise_times, int num_times,
                       const double* means_dev, int64_t mean_stride, const double* ratio_dev,
                       double rcond, uint64_t skip_below, double* value_sums_dev,
                       double* beta_dev, double* history_dev, cudaStream_t stream) {
  (void)mean_stride;     // one payoff: the means of payoff 0
  if (!lsm_persistent_ok(h)) {
    set_error("the persistent backward induction does not apply to this problem "
              "(tqf_lsm_persistent_eligible)");
    return TQF_ERR_UNSUPPORTED;
  }
  if (num_times > h->times_cap) {
    void* old = h->times_dev;
    dev_release(&old, 1);
    h->times_dev = nullptr;
    h->times_cap = 0;
    void* fresh = nullptr;
    const int rc = dev_alloc(&fresh, sizeof(int) * num_times);
    if (rc != TQF_OK) return rc;
    h->times_dev = static_cast<int*>(fresh);
    h->times_cap = num_times;
  }
  TQF_CUDA_OK(cudaMemcpyAsync(h->times_dev, exercise_times, sizeof(int) * num_times,
                              cudaMemcpyHostToDevice, stream));
  TQF_CUDA_OK(cudaMemsetAsync(h->ctrl_dev, 0, 4 * sizeof(unsigned long long), stream));
  const size_t need = static_cast<size_t>(kSMs + 8) * kLsmFastNS + 64;
  if (need > h->partials_doubles) {
    if (h->external_partials) {
      set_error("caller-provided LSM partials workspace is too small");
      return TQF_ERR_INVALID_ARGUMENT;
    }
    void* old = h->partials_dev;
    dev_release(&old, 1);
    h->partials_dev = nullptr;
    h->partials_doubles = 0;
    void* fresh = nullptr;
    const int rc = dev_alloc(&fresh, need * sizeof(double));
    if (rc != TQF_OK) return rc;
    h->partials_dev = static_cast<double*>(fresh);
    h->partials_doubles = need;
  }
  const int rc = h->desc.dtype == TQF_F64
                     ? run_persistent_t<double>(h, num_times, means_dev, ratio_dev, rcond, skip_below,
                                                value_sums_dev, beta_dev, history_dev, stream)
                     : run_persistent_t<float>(h, num_times, means_dev, ratio_dev, rcond, skip_below,
                                               value_sums_dev, beta_dev, history_dev, stream);
  if (rc == TQF_OK && h->peer_world > 1) h->peer_epoch += static_cast<unsigned long long>(num_times);
  return rc;
}

}  // namespace tqf
