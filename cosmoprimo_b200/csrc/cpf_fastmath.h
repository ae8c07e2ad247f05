// cpf_fastmath.h — fp64 log and exp for the Wallish2018 kernel's log(k P) and exp(.)/k (bao_filter.py:371, 413), which were 42 % of its
// instructions with the CUDA library functions (98 and 58 instructions per call with their special-case handling inlined 56 times).
// Both are accurate to ~1 ulp on their fast path (normal positive arguments for log, |x| < 700 for exp) and hand everything else (zero,
// denormal, negative, Inf, NaN, overflow) to the library function, so special values behave as in numpy.  Host + device: tests/emul
// runs the same code on the CPU against libm.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CPF_FHD __host__ __device__ __forceinline__
#else
#define CPF_FHD inline
#endif

namespace cpf {

CPF_FHD int fm_hi(const double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
CPF_FHD int fm_lo(const double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u;
#endif
}
CPF_FHD double fm_make(const int hi, const int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}
// reciprocal seed with ~20 good bits (MUFU.RCP64H); two Newton steps in the caller make it exact to rounding
CPF_FHD double fm_rcp_seed(const double g) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(g));
  return r;
#else
  return (double)(float)(1. / g);
#endif
}

// Polynomial coefficients.  On the device they live in constant memory so that every Horner step is ONE DFMA with a constant-bank operand:
// as literals the compiler rebuilt each 64-bit constant with two moves in front of every use (no registers to spare under the kernel's
// 128-register cap), which tripled the instruction count of the polynomials.
#define CPF_FM_LOG_COEFFS {2. / 3., 2. / 5., 2. / 7., 2. / 9., 2. / 11., 2. / 13., 2. / 15., 2. / 17., 2. / 19.}
#define CPF_FM_EXP_COEFFS {1., 1., 0.5, 1. / 6., 1. / 24., 1. / 120., 1. / 720., 1. / 5040., 1. / 40320., 1. / 362880., 1. / 3628800., 1. / 39916800., 1. / 479001600., 1. / 6227020800.}
#define CPF_FM_MISC {0.693147180369123816490, 1.90821492927058770002e-10, 1.4426950408889634074, 6755399441055744.}   /* ln2_hi, ln2_lo (fdlibm), 1/ln 2, 1.5 * 2^52 */
// 10^r = sum_k (ln 10)^k r^k / k!, |r| <= log10(2)/2 = 0.1505 (term 14: 4e-18)
#define CPF_FM_EXP10_COEFFS {1.000000000000000000000e+00, 2.302585092994045901094e+00, 2.650949055239199214640e+00, 2.034678592293476029340e+00, 1.171255148912266896843e+00, 5.393829291955813953763e-01, 2.069958486968681010687e-01, 6.808936507443706653842e-02, 1.959769462647852414361e-02, 5.013928833775440144227e-03, 1.154499778998434788777e-03, 2.416667255442469424212e-04, 4.637151664257219596706e-05, 8.213412535439386743565e-06}
#define CPF_FM_MISC10 {3.010299955494701862335e-01, 1.145110089802183842107e-10, 3.321928094887362181709e+00, 4.342944819032518166679e-01, 1.098319650216765072739e-17}   /* log10(2) hi (low 21 bits clear), lo ; log2(10) ; log10(e) hi, lo */
#if defined(__CUDACC__)
static __constant__ double kFmExp10C[14] = CPF_FM_EXP10_COEFFS;
static __constant__ double kFmMisc10[5] = CPF_FM_MISC10;
#endif
static const double kFmExp10H[14] = CPF_FM_EXP10_COEFFS;
static const double kFmMisc10H[5] = CPF_FM_MISC10;
#if defined(__CUDA_ARCH__)
#define CPF_FM_EXP10(i) kFmExp10C[i]
#define CPF_FM_MISC10V(i) kFmMisc10[i]
#else
#define CPF_FM_EXP10(i) kFmExp10H[i]
#define CPF_FM_MISC10V(i) kFmMisc10H[i]
#endif
#if defined(__CUDACC__)
static __constant__ double kFmLogC[9] = CPF_FM_LOG_COEFFS;
static __constant__ double kFmExpC[14] = CPF_FM_EXP_COEFFS;
static __constant__ double kFmMisc[4] = CPF_FM_MISC;
#endif
static const double kFmLogH[9] = CPF_FM_LOG_COEFFS;
static const double kFmExpH[14] = CPF_FM_EXP_COEFFS;
static const double kFmMiscH[4] = CPF_FM_MISC;
#if defined(__CUDA_ARCH__)
#define CPF_FM_LOG(i) kFmLogC[i]
#define CPF_FM_EXP(i) kFmExpC[i]
#define CPF_FM_MISCV(i) kFmMisc[i]
#else
#define CPF_FM_LOG(i) kFmLogH[i]
#define CPF_FM_EXP(i) kFmExpH[i]
#define CPF_FM_MISCV(i) kFmMiscH[i]
#endif

// natural logarithm.  x = 2^e m, m in [sqrt 1/2, sqrt 2); log m = 2 atanh(s), s = (m - 1)/(m + 1), |s| <= 0.1716:
// 2 s + 2 s^3/3 + ... + 2 s^19/19 (next term 8e-18); log x = e ln2_hi + (log m + e ln2_lo).
// fast path of the logarithms: x = 2^e m (normal, positive), returns log m and e
CPF_FHD double fm_log_core(const double x, int& e) {
  int hi = fm_hi(x);
  const int lo = fm_lo(x);
  e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; ++e; }                        // m >= sqrt 2 (to 20 bits): halve it
  const double m = fm_make(hi, lo);
  const double f = m - 1., g = m + 1.;
  double r = fm_rcp_seed(g);
  r = fma(fma(-g, r, 1.), r, r);
  r = fma(fma(-g, r, 1.), r, r);
  double s = f * r;
  s = fma(fma(-g, s, f), r, s);
  const double z = s * s;
  double p = CPF_FM_LOG(8);
#pragma unroll
  for (int i = 7; i >= 0; --i) p = fma(p, z, CPF_FM_LOG(i));
  return fma(s * z, p, s + s);
}
CPF_FHD bool fm_log_fast_path(const double x) { return (unsigned)(fm_hi(x) - 0x00100000) < 0x7fe00000u; }   // not zero, denormal, negative, Inf, NaN

CPF_FHD double fast_log(const double x) {
  if (!fm_log_fast_path(x)) return log(x);
  int e;
  const double lm = fm_log_core(x, e);
  const double ed = (double)e;
  return fma(ed, CPF_FM_MISCV(0), fma(ed, CPF_FM_MISCV(1), lm));
}

// decimal logarithm (the spline kernels' log10 of a table, interpolator.py:42-87 / jax.py:153), < 1.5 ulp on the fast path: log m = 2 s + tail with
// the rounding residual of the quotient s kept in the tail, times log10(e) as a two-word product, plus e log10(2) in two words
CPF_FHD double fast_log10(const double x) {
  if (!fm_log_fast_path(x)) return log10(x);
  int hi = fm_hi(x);
  const int lo = fm_lo(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; ++e; }
  const double m = fm_make(hi, lo);
  const double f = m - 1., g = m + 1.;
  double r = fm_rcp_seed(g);
  r = fma(fma(-g, r, 1.), r, r);
  r = fma(fma(-g, r, 1.), r, r);
  double s = f * r;
  s = fma(fma(-g, s, f), r, s);
  const double s_lo = fma(-g, s, f) * r;                                   // f / g = s + s_lo
  const double z = s * s;
  double p = CPF_FM_LOG(8);
#pragma unroll
  for (int i = 7; i >= 0; --i) p = fma(p, z, CPF_FM_LOG(i));
  const double two_s = s + s;
  const double tail = fma(s * z, p, s_lo + s_lo);                          // log m = two_s + tail
  const double l10e = CPF_FM_MISC10V(3);
  const double ph = two_s * l10e;
  const double pl = fma(two_s, l10e, -ph) + fma(two_s, CPF_FM_MISC10V(4), tail * l10e);
  const double ed = (double)e;
  return fma(ed, CPF_FM_MISC10V(0), ph + fma(ed, CPF_FM_MISC10V(1), pl));
}

// exponential: n = rint(x / ln 2), r = x - n ln 2 (|r| <= 0.3466), exp r by its Taylor series to r^13/13! (next term 4e-18), times 2^n
CPF_FHD double fast_exp(const double x) {
  if (!(fabs(x) < 700.)) return exp(x);                                   // overflow / underflow range, Inf, NaN
  const double magic = CPF_FM_MISCV(3);                                    // 1.5 * 2^52: the low word of x/ln2 + magic is n
  const double t = fma(x, CPF_FM_MISCV(2), magic);
  const int n = fm_lo(t);
  const double nd = t - magic;
  double r = fma(-nd, CPF_FM_MISCV(0), x);
  r = fma(-nd, CPF_FM_MISCV(1), r);
  double p = CPF_FM_EXP(13);
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, CPF_FM_EXP(i));
  return fm_make(fm_hi(p) + (n << 20), fm_lo(p));                         // p in [0.7, 1.42], |n| <= 1010: the result is normal
}

// 10^x (the spline kernels' 10**tmp, jax.py:191): n = rint(x log2 10), r = x - n log10 2 (|r| <= 0.1505), 10^r by its series, times 2^n; < 2 ulp
CPF_FHD double fast_exp10(const double x) {
  if (!(fabs(x) < 300.)) {                                                // overflow / underflow range, Inf, NaN
#if defined(__CUDA_ARCH__)
    return exp10(x);
#else
    return pow(10., x);
#endif
  }
  const double magic = CPF_FM_MISCV(3);
  const double t = fma(x, CPF_FM_MISC10V(2), magic);
  const int n = fm_lo(t);
  const double nd = t - magic;
  double r = fma(-nd, CPF_FM_MISC10V(0), x);                              // exact: |n| < 1024 and the constant has 21 trailing zero bits
  r = fma(-nd, CPF_FM_MISC10V(1), r);
  double p = CPF_FM_EXP10(13);
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, CPF_FM_EXP10(i));
  return fm_make(fm_hi(p) + (n << 20), fm_lo(p));                         // p in [0.70, 1.42], |n| <= 997: the result is normal
}

}  // namespace cpf
