// CPU check of cpf_fastmath.h (the same code the kernels run) against long-double libm: fast_log, fast_exp (Wallish2018 kernel),
// fast_log10, fast_exp10 (spline kernels).  Prints the largest error in ulps of the result and "OK" when all are < 2 ulp and the
// special values go where libm sends them.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>
#include "../../cosmoprimo_b200/csrc/cpf_fastmath.h"

static double ulps(const double got, const long double want) {
  if (want == 0.L) return got == 0. ? 0. : 1e9;
  int ex;
  frexpl(want, &ex);
  const long double ulp = ldexpl(1.L, ex - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}

int main() {
  using namespace cpf;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> u01(0., 1.);
  double e_log = 0., e_log10 = 0., e_exp = 0., e_exp10 = 0.;
  for (int i = 0; i < 2000000; ++i) {
    // arguments: wide log-uniform, and dense around 1 where the logarithm cancels
    const double x = i % 3 == 0 ? 1. + (u01(rng) - 0.5) * 0.6 : std::pow(10., -300. + 600. * u01(rng));
    e_log = std::fmax(e_log, ulps(fast_log(x), logl((long double)x)));
    e_log10 = std::fmax(e_log10, ulps(fast_log10(x), log10l((long double)x)));
    const double y = i % 3 == 0 ? (u01(rng) - 0.5) * 2. : (u01(rng) - 0.5) * 1390.;
    e_exp = std::fmax(e_exp, ulps(fast_exp(y), expl((long double)y)));
    const double z = i % 3 == 0 ? (u01(rng) - 0.5) * 2. : (u01(rng) - 0.5) * 598.;
    e_exp10 = std::fmax(e_exp10, ulps(fast_exp10(z), powl(10.L, (long double)z)));
  }
  std::printf("max error (ulp): fast_log %.3f, fast_log10 %.3f, fast_exp %.3f, fast_exp10 %.3f\n", e_log, e_log10, e_exp, e_exp10);
  bool ok = e_log < 2. && e_log10 < 2. && e_exp < 2. && e_exp10 < 2.;
  // exact powers, special values
  for (int k = -300; k <= 300; ++k) {
    const double p = fast_exp10((double)k);
    if (ulps(p, powl(10.L, (long double)k)) >= 2.) { std::printf("10^%d off\n", k); ok = false; }
  }
  ok = ok && fast_log10(1.) == 0. && fast_log10(0.) == -INFINITY && std::isnan(fast_log10(-1.)) && std::isnan(fast_log10(NAN)) && fast_log10(INFINITY) == INFINITY;
  ok = ok && fast_exp10(0.) == 1. && fast_exp10(-INFINITY) == 0. && fast_exp10(INFINITY) == INFINITY && std::isnan(fast_exp10(NAN)) && fast_exp10(400.) == INFINITY && fast_exp10(-400.) == 0.;
  ok = ok && std::fabs(fast_log10(5e-324) - log10(5e-324)) < 1e-12 && std::fabs(fast_exp10(-310.) / 1e-310 - 1.) < 1e-6;
  std::printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
