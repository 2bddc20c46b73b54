// Hand-written FP64 device math for the fused path kernels.
//
// The fused mode is bound by the FP64 pipe (64 DFMA/clk/SM) and by issue slots,
// so the transcendental functions behind every normal draw are written for a
// minimal DFMA count with (almost) no integer / branch overhead, instead of
// calling libdevice (normcdfinv, log, sincos, sqrt carry special-case handling
// that more than doubles the issued instructions -- profiles/README.md).
//
// Everything is vectorised over K independent arguments held by ONE thread:
// sm_100a's DFMA cannot take a constant-bank operand, so every polynomial
// coefficient costs a load (LDC.128 fetches two); evaluating K Horner chains
// side by side amortises that load over K DFMAs and gives the FP64 pipe K
// independent dependency chains per thread.
//
// Coefficients: tools/fit_math.py -> tqf_math_coef.h (highest order first).
// Accuracy (checked on the GPU against mpmath by tests/test_gpu_math.py):
// <= 3 ulp for log, <= 1 ulp for sqrt, <= 2 ulp for sincos on their stated domains, <= 5e-16 relative
// for ndtri.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tqf_math_coef.h"

namespace tqf {
namespace fm {

// Polynomial coefficients always come from the constant bank (TQF_COEF): a
// volatile `ld.const.v2.f64` becomes an LDCU.128 into uniform registers next to
// its DFMAs, which then take the coefficient as a UR operand -- measured
// 98-100% of the DFMA peak against 81-88% for LDS-fed register operands
// (tools/microbench/dfma_bench.cu).
//
// The "Tab" objects say where the {T, 1/c} table of the table logarithm
// (tqf_logtab.inc, 20 KB) is read from:
//  * SmemTab: a per-CTA copy in shared memory (fused path kernels; one LDS.128
//    per draw, divergent index);
//  * ConstTab: the device-global copy through the read-only path (fill kernels,
//    test hook, kernels that cannot spare the shared memory).
struct SmemTab {
  uint32_t logbase;  // shared-window address of the log table (0: none)
  __device__ __forceinline__ SmemTab() : logbase(0) {}
  __device__ __forceinline__ explicit SmemTab(const double* s_logtab)
      : logbase(static_cast<uint32_t>(__cvta_generic_to_shared(s_logtab))) {}
  // entry at byte offset `off16` (a multiple of 16 below 16 * TQF_LOGTAB_COUNT)
  __device__ __forceinline__ void log_entry(uint32_t off16, double* t, double* s) const {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(*t), "=d"(*s) : "r"(logbase + off16));
  }
};
struct ConstTab {
  const double* logtab;  // device-global table
  __device__ __forceinline__ ConstTab() : logtab(nullptr) {}
  __device__ __forceinline__ explicit ConstTab(const double* g) : logtab(g) {}
  __device__ __forceinline__ void log_entry(uint32_t off16, double* t, double* s) const {
    const double2 e = __ldg(reinterpret_cast<const double2*>(
        reinterpret_cast<const char*>(logtab) + off16));
    *t = e.x;
    *s = e.y;
  }
};
// Copies the log table into shared memory (call before a __syncthreads()).
__device__ __forceinline__ void fill_smem_logtab(double* dst, const double* src, int tid,
                                                 int nthreads) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int i = tid; i < TQF_LOGTAB_COUNT; i += nthreads) d2[i] = __ldg(s2 + i);
}

// ---- Horner steps as single asm statements --------------------------------
// One `asm volatile` holds the coefficient-pair load and the K fused
// multiply-adds that consume it.  Volatile statements keep their order, so the
// K chains advance in lock step (K independent DFMAs between dependent ones)
// and the coefficient is fetched by LDCU into a uniform register that the
// DFMAs read as a UR operand -- the only shape that reaches the FP64 peak
// (tools/microbench/dfma_bench.cu: 98-100% vs 81-90% for register operands).
#define TQF_CADDR(off) "l"(__cvta_generic_to_constant(TQF_COEF + (off)))

template <int K>
struct HornerAsm;

template <>
struct HornerAsm<8> {
  // p = c0 * y + c1
  static __device__ __forceinline__ void first(int off, const double (&y)[8], double (&p)[8]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%16];\n\t"
        "fma.rn.f64 %0, c0, %8, c1;\n\tfma.rn.f64 %1, c0, %9, c1;\n\t"
        "fma.rn.f64 %2, c0, %10, c1;\n\tfma.rn.f64 %3, c0, %11, c1;\n\t"
        "fma.rn.f64 %4, c0, %12, c1;\n\tfma.rn.f64 %5, c0, %13, c1;\n\t"
        "fma.rn.f64 %6, c0, %14, c1;\n\tfma.rn.f64 %7, c0, %15, c1;\n\t}"
        : "=d"(p[0]), "=d"(p[1]), "=d"(p[2]), "=d"(p[3]), "=d"(p[4]), "=d"(p[5]), "=d"(p[6]),
          "=d"(p[7])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), "d"(y[4]), "d"(y[5]), "d"(y[6]), "d"(y[7]),
          TQF_CADDR(off));
  }
  // p = (p * y + c0) * y + c1
  static __device__ __forceinline__ void pair(int off, const double (&y)[8], double (&p)[8]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%16];\n\t"
        "fma.rn.f64 %0, %0, %8, c0;\n\tfma.rn.f64 %1, %1, %9, c0;\n\t"
        "fma.rn.f64 %2, %2, %10, c0;\n\tfma.rn.f64 %3, %3, %11, c0;\n\t"
        "fma.rn.f64 %4, %4, %12, c0;\n\tfma.rn.f64 %5, %5, %13, c0;\n\t"
        "fma.rn.f64 %6, %6, %14, c0;\n\tfma.rn.f64 %7, %7, %15, c0;\n\t"
        "fma.rn.f64 %0, %0, %8, c1;\n\tfma.rn.f64 %1, %1, %9, c1;\n\t"
        "fma.rn.f64 %2, %2, %10, c1;\n\tfma.rn.f64 %3, %3, %11, c1;\n\t"
        "fma.rn.f64 %4, %4, %12, c1;\n\tfma.rn.f64 %5, %5, %13, c1;\n\t"
        "fma.rn.f64 %6, %6, %14, c1;\n\tfma.rn.f64 %7, %7, %15, c1;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3]), "+d"(p[4]), "+d"(p[5]), "+d"(p[6]),
          "+d"(p[7])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), "d"(y[4]), "d"(y[5]), "d"(y[6]), "d"(y[7]),
          TQF_CADDR(off));
  }
  // p = p * y + c0
  static __device__ __forceinline__ void single(int off, const double (&y)[8], double (&p)[8]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%16];\n\t"
        "fma.rn.f64 %0, %0, %8, c0;\n\tfma.rn.f64 %1, %1, %9, c0;\n\t"
        "fma.rn.f64 %2, %2, %10, c0;\n\tfma.rn.f64 %3, %3, %11, c0;\n\t"
        "fma.rn.f64 %4, %4, %12, c0;\n\tfma.rn.f64 %5, %5, %13, c0;\n\t"
        "fma.rn.f64 %6, %6, %14, c0;\n\tfma.rn.f64 %7, %7, %15, c0;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3]), "+d"(p[4]), "+d"(p[5]), "+d"(p[6]),
          "+d"(p[7])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), "d"(y[4]), "d"(y[5]), "d"(y[6]), "d"(y[7]),
          TQF_CADDR(off));
  }
};

template <>
struct HornerAsm<4> {
  static __device__ __forceinline__ void first(int off, const double (&y)[4], double (&p)[4]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%8];\n\t"
        "fma.rn.f64 %0, c0, %4, c1;\n\tfma.rn.f64 %1, c0, %5, c1;\n\t"
        "fma.rn.f64 %2, c0, %6, c1;\n\tfma.rn.f64 %3, c0, %7, c1;\n\t}"
        : "=d"(p[0]), "=d"(p[1]), "=d"(p[2]), "=d"(p[3])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), TQF_CADDR(off));
  }
  static __device__ __forceinline__ void pair(int off, const double (&y)[4], double (&p)[4]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%8];\n\t"
        "fma.rn.f64 %0, %0, %4, c0;\n\tfma.rn.f64 %1, %1, %5, c0;\n\t"
        "fma.rn.f64 %2, %2, %6, c0;\n\tfma.rn.f64 %3, %3, %7, c0;\n\t"
        "fma.rn.f64 %0, %0, %4, c1;\n\tfma.rn.f64 %1, %1, %5, c1;\n\t"
        "fma.rn.f64 %2, %2, %6, c1;\n\tfma.rn.f64 %3, %3, %7, c1;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), TQF_CADDR(off));
  }
  static __device__ __forceinline__ void single(int off, const double (&y)[4], double (&p)[4]) {
    asm volatile(
        "{\n\t.reg .f64 c0, c1;\n\t"
        "ld.const.v2.f64 {c0, c1}, [%8];\n\t"
        "fma.rn.f64 %0, %0, %4, c0;\n\tfma.rn.f64 %1, %1, %5, c0;\n\t"
        "fma.rn.f64 %2, %2, %6, c0;\n\tfma.rn.f64 %3, %3, %7, c0;\n\t}"
        : "+d"(p[0]), "+d"(p[1]), "+d"(p[2]), "+d"(p[3])
        : "d"(y[0]), "d"(y[1]), "d"(y[2]), "d"(y[3]), TQF_CADDR(off));
  }
};

// Generic K (1, 2, ...): plain C++ with the same volatile coefficient loads.
template <int K>
struct HornerAsm {
  static __device__ __forceinline__ void ld(int off, double* c0, double* c1) {
    asm volatile("ld.const.v2.f64 {%0, %1}, [%2];" : "=d"(*c0), "=d"(*c1) : TQF_CADDR(off));
  }
  static __device__ __forceinline__ void first(int off, const double (&y)[K], double (&p)[K]) {
    double c0, c1;
    ld(off, &c0, &c1);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(c0, y[k], c1);
  }
  static __device__ __forceinline__ void pair(int off, const double (&y)[K], double (&p)[K]) {
    double c0, c1;
    ld(off, &c0, &c1);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(fma(p[k], y[k], c0), y[k], c1);
  }
  static __device__ __forceinline__ void single(int off, const double (&y)[K], double (&p)[K]) {
    double c0, c1;
    ld(off, &c0, &c1);
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(p[k], y[k], c0);
  }
};

// p[k] = sum_i c[OFF+i] y[k]^(N-1-i)  (highest order first, padded to even).
template <int OFF, int N, int K, class Tab>
__device__ __forceinline__ void horner_v(const Tab&, const double (&y)[K], double (&p)[K]) {
  HornerAsm<K>::first(OFF, y, p);
#pragma unroll
  for (int i = 2; i + 1 < N; i += 2) HornerAsm<K>::pair(OFF + i, y, p);
  if (N & 1) HornerAsm<K>::single(OFF + N - 1, y, p);
}

// 1/d, d normal and positive: MUFU.RCP64H (2^-23) + two Newton steps.
__device__ __forceinline__ double rcp_pos(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// sqrt(v), v normal and positive: MUFU.RSQ64H (y0 = rsqrt to 2^-22) and two
// Newton steps on g ~ sqrt(v) that share h = y0 / 2 (an exponent decrement on
// the integer pipe):  g <- g + (v - g^2) h.  The first step leaves 1.5 * 2^-44,
// the second 2^-44 * 2^-22 (h is only 22 bits good, which is enough for a
// correction term).  5 FP64 operations; <= 1 ulp (tests/test_gpu_math.py).
__device__ __forceinline__ double sqrt_pos(double v) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  double g = v * y;
  g = fma(fma(-g, g, v), h, g);
  g = fma(fma(-g, g, v), h, g);
  return g;
}

// log(a[k]), a normal, positive and finite (no zero / inf / nan / denormal
// handling: the callers' arguments are in [2^-60, 2]).
//   a = 2^k m, m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1),
//   log a = k ln2 + 2 s + s^3 R(s^2).
template <int K, class Tab>
__device__ __forceinline__ void log_pos_v(const Tab& tab, const double (&a)[K], double (&out)[K]) {
  const double kLn2Hi = 6.93147180369123816490e-01;  // fdlibm split
  const double kLn2Lo = 1.90821492927058770002e-10;
  double s[K], z[K], kd[K], R[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int hi = __double2hiint(a[k]);
    const int lo = __double2loint(a[k]);
    const int e = (hi - 0x3fe6a09e) >> 20;
    hi -= e << 20;
    const double m = __hiloint2double(hi, lo);
    kd[k] = static_cast<double>(e);
    s[k] = (m - 1.0) * rcp_pos(m + 1.0);
    z[k] = s[k] * s[k];
  }
  horner_v<TQF_LOG_R_OFF, TQF_LOG_R_N, K>(tab, z, R);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // k ln2_hi + (s (2 + z R) + k ln2_lo)
    const double q = fma(z[k], R[k], 2.0);
    out[k] = fma(kd[k], kLn2Hi, fma(s[k], q, kd[k] * kLn2Lo));
  }
}

// y[k] = -log(a[k]) - TQF_NDTRI_C_MID for a in [2^-15, 1] by table lookup:
//   a = 2^e m;  entry i = bits [13, 24) of the high word of a (low 4 exponent
//   bits, top 7 mantissa bits: 2048 entries, any bit pattern stays in bounds);
//   r = a * s[i] - 1 with s = 2^-e / c (one exact FMA, |r| <= 2^-8);
//   y = T[i] - log1p(r),  T[i] = -MID - log(2^e c).
// 8 FP64 instructions, 2 integer instructions and one 16-byte table load,
// against 18 FP64 + MUFU + I2F + 6 integer for log_pos_v.  Absolute error
// <= 1e-15 (T and T - r are each rounded once at magnitude <= 8), which is what
// the polynomial in w needs: d(ndtri)/ndtri = (P'/P) dw <= 0.26 dw.  Arguments
// below 2^-15 alias another row and give garbage: the caller recomputes those
// (far tail) draws with log_pos_v.
template <int K, class Tab>
__device__ __forceinline__ void neg_log_mid_tab_v(const Tab& tab, const double (&a)[K],
                                                  double (&y)[K], uint32_t* hmin_out) {
  double r[K], tr[K], q[K];
  uint32_t hmin = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t uh = static_cast<uint32_t>(__double2hiint(a[k]));
    hmin = uh < hmin ? uh : hmin;
    const uint32_t off16 = (uh >> (20 - TQF_LOGTAB_BITS - 4)) & ((TQF_LOGTAB_COUNT - 1) << 4);
    double t, sc;
    tab.log_entry(off16, &t, &sc);
    r[k] = fma(a[k], sc, -1.0);
    tr[k] = t - r[k];
  }
  horner_v<TQF_LOG1P_OFF, TQF_LOG1P_N, K>(tab, r, q);
#pragma unroll
  for (int k = 0; k < K; ++k) y[k] = fma(r[k] * r[k], q[k], tr[k]);
  *hmin_out = hmin;
}

// Tail polynomial of ndtri (TQF_NDTRI_T, degree 24) for ONE argument: the
// branch is entered by single lanes, so its cost is dependent-instruction
// latency -- even and odd powers run as two interleaved Horner chains in y^2
// (13 dependent FMAs instead of 24); each ld.const pair feeds both chains.
__device__ __forceinline__ double ndtri_tail_poly(double y) {
  static_assert(TQF_NDTRI_T_N == 25, "tail polynomial layout");
  const double y2 = y * y;
  double e, o, c0, c1;
  HornerAsm<1>::ld(TQF_NDTRI_T_OFF, &e, &o);
#pragma unroll
  for (int i = 2; i < TQF_NDTRI_T_N - 1; i += 2) {
    HornerAsm<1>::ld(TQF_NDTRI_T_OFF + i, &c0, &c1);
    e = fma(e, y2, c0);
    o = fma(o, y2, c1);
  }
  HornerAsm<1>::ld(TQF_NDTRI_T_OFF + TQF_NDTRI_T_N - 1, &c0, &c1);
  e = fma(e, y2, c0);
  return fma(y, o, e);
}

// Inverse normal CDF of u = (1 + t) / 2 (t = 2u - 1 exact), |t| < 1:
//   sqrt(2) erfinv(t) = t * P(w),  w = -log(1 - t^2)
// (the parametrisation of M. Giles, "Approximating the erfinv function", 2010,
// with our own double-precision fits).  The central branch w < 6.25 covers
// |t| < 0.99806, i.e. 99.8% of uniform draws, and takes w from the table
// logarithm; the tail branch is entered by a thread only when one of its K
// draws needs it and recomputes w with the full-precision logarithm.
template <int K, class Tab>
__device__ __forceinline__ void ndtri_t_v(const Tab& tab, const double (&t)[K], double (&zout)[K]) {
  double a[K], y[K], p[K];
#pragma unroll
  for (int k = 0; k < K; ++k) a[k] = fma(-t[k], t[k], 1.0);
  uint32_t hmin;
  neg_log_mid_tab_v<K>(tab, a, y, &hmin);
  horner_v<TQF_NDTRI_C_OFF, TQF_NDTRI_C_N, K>(tab, y, p);
  // positive doubles order like their high words: one integer min + compare
  // finds out whether any draw has 1 - t^2 < exp(-6.25)
  if (hmin < TQF_NDTRI_TAIL_HI) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint32_t hk = static_cast<uint32_t>(__double2hiint(a[k]));
      if (hk < TQF_NDTRI_TAIL_HI) {
        // w: from the table value while the table covers a, else the full logarithm
        double w = y[k] + TQF_NDTRI_C_MID;
        if (hk < ((1023u + TQF_LOGTAB_EMIN) << 20)) {
          const double ak[1] = {a[k]};
          double lg[1];
          log_pos_v<1>(tab, ak, lg);
          w = -lg[0];
        }
        p[k] = ndtri_tail_poly(sqrt_pos(w) - TQF_NDTRI_T_MID);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) zout[k] = t[k] * p[k];
}

// sin(v[k]), cos(v[k]) for v in [0, 2 pi] (Box-Muller angle): Cody-Waite
// reduction by pi/2 with an FMA pair, then the kernels on |f| <= pi/4.
template <int K, class Tab>
__device__ __forceinline__ void sincos_2pi_v(const Tab& tab, const double (&v)[K], double (&sn)[K],
                                             double (&cs)[K]) {
  const double kTwoOverPi = 6.36619772367581382433e-01;
  const double kPio2Hi = 1.57079632679489655800e+00;
  const double kPio2Lo = 6.12323399573676603587e-17;
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  double f[K], z[K], ps[K], pc[K];
  int j[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double jm = fma(v[k], kTwoOverPi, kMagic);
    j[k] = __double2loint(jm);
    const double jd = jm - kMagic;
    f[k] = fma(-jd, kPio2Lo, fma(-jd, kPio2Hi, v[k]));
    z[k] = f[k] * f[k];
  }
  horner_v<TQF_SIN_P_OFF, TQF_SIN_P_N, K>(tab, z, ps);
  horner_v<TQF_COS_P_OFF, TQF_COS_P_N, K>(tab, z, pc);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = f[k] * ps[k];
    const double c = pc[k];
    double rs = (j[k] & 1) ? c : s;
    double rc = (j[k] & 1) ? s : c;
    if (j[k] & 2) rs = -rs;
    if ((j[k] + 1) & 2) rc = -rc;
    sn[k] = rs;
    cs[k] = rc;
  }
}

// float32 inverse CDF: sqrt(2) erfinv(t) = t P(w), w = -log(1 - t^2), with the
// SFU logarithm and float coefficients as FFMA immediates (~18 instructions).
// t = +-1 (u rounded to 0 or 1) gives +-inf like the reference.
__device__ __forceinline__ float ndtri_t_f32(float t) {
  const float a = fmaf(-t, t, 1.0f);
  const float w = -__logf(a);
  float p;
  if (w < 6.25f) {
    const float c[TQF_NDTRI_F32_C_N] = {TQF_NDTRI_F32_C_LIST};
    const float y = w - TQF_NDTRI_F32_C_MID;
    p = c[0];
#pragma unroll
    for (int i = 1; i < TQF_NDTRI_F32_C_N; ++i) p = fmaf(p, y, c[i]);
  } else {
    const float c[TQF_NDTRI_F32_T_N] = {TQF_NDTRI_F32_T_LIST};
    const float y = sqrtf(w) - TQF_NDTRI_F32_T_MID;
    p = c[0];
#pragma unroll
    for (int i = 1; i < TQF_NDTRI_F32_T_N; ++i) p = fmaf(p, y, c[i]);
  }
  return t * p;
}

// scalar conveniences (fill kernels, test hook): coefficients straight from
// the constant bank.
__device__ __forceinline__ double log_pos(double a) {
  const double in[1] = {a};
  double out[1];
  log_pos_v<1>(ConstTab(), in, out);
  return out[0];
}
// ndtri(u) for u = 1/2 + q, q exact; `logtab` = device-global log table.
__device__ __forceinline__ double ndtri_q(double q, const double* logtab) {
  const double in[1] = {q + q};
  double out[1];
  ndtri_t_v<1>(ConstTab(logtab), in, out);
  return out[0];
}
__device__ __forceinline__ void sincos_2pi(double v, double* sn, double* cs) {
  const double in[1] = {v};
  double s[1], c[1];
  sincos_2pi_v<1>(ConstTab(), in, s, c);
  *sn = s[0];
  *cs = c[0];
}

}  // namespace fm
}  // namespace tqf
