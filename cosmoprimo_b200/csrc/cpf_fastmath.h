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
#if defined(__CUDACC__)
__constant__ double kFmLogC[9] = CPF_FM_LOG_COEFFS;
__constant__ double kFmExpC[14] = CPF_FM_EXP_COEFFS;
__constant__ double kFmMisc[4] = CPF_FM_MISC;
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
CPF_FHD double fast_log(const double x) {
  int hi = fm_hi(x);
  const int lo = fm_lo(x);
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(x);          // zero, denormal, negative, Inf, NaN
  int e = (hi >> 20) - 1023;
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
  const double lm = fma(s * z, p, s + s);
  const double ed = (double)e;
  return fma(ed, CPF_FM_MISCV(0), fma(ed, CPF_FM_MISCV(1), lm));
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

}  // namespace cpf
