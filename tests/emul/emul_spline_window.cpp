// CPU check of spline_window_weights (csrc/cpf_spline_core.h), the weights behind cpf_spline_eval_rows: the windowed
// weighted sum must reproduce the full natural / clamped cubic spline (long-double Thomas solve of scipy's slope system,
// scipy/interpolate/_cubic.py) on a log-spaced grid used in LINEAR abscissa (integrate_sigma_r2, interpolator.py:289),
// for steep power-law-like ordinates, at interior, edge and end-interval queries.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../cosmoprimo_b200/csrc/cpf_spline_core.h"

using namespace cpf;
typedef long double ld;

static std::vector<ld> full_slopes(const std::vector<double>& x, const std::vector<double>& y, int bc) {
  const int n = (int)x.size();
  std::vector<ld> lo(n), di(n), up(n), rhs(n), s(n);
  for (int i = 0; i < n; ++i) {
    double l, d, u;
    spline_row(x.data(), n, bc, i, l, d, u);
    lo[i] = l; di[i] = d; up[i] = u;
    if (i == 0) rhs[i] = bc == 1 ? 0 : 3 * ((ld)y[1] - y[0]);
    else if (i == n - 1) rhs[i] = bc == 1 ? 0 : 3 * ((ld)y[n - 1] - y[n - 2]);
    else {
      const ld dm = (ld)x[i] - x[i - 1], dp = (ld)x[i + 1] - x[i];
      rhs[i] = 3 * (dp * (((ld)y[i] - y[i - 1]) / dm) + dm * (((ld)y[i + 1] - y[i]) / dp));
    }
  }
  for (int i = 1; i < n; ++i) { const ld m = lo[i] / di[i - 1]; di[i] -= m * up[i - 1]; rhs[i] -= m * rhs[i - 1]; }
  s[n - 1] = rhs[n - 1] / di[n - 1];
  for (int i = n - 2; i >= 0; --i) s[i] = (rhs[i] - up[i] * s[i + 1]) / di[i];
  return s;
}

static ld full_eval(const std::vector<double>& x, const std::vector<double>& y, const std::vector<ld>& s, double xv, ld* mag) {
  // Hermite form (the power basis around x_i cancels catastrophically when the ordinates span many decades per interval)
  const int i = spline_interval(x.data(), (int)x.size(), xv);
  const ld dx = (ld)x[i + 1] - x[i], u = ((ld)xv - x[i]) / dx;
  const ld t0 = (1 + 2 * u) * (1 - u) * (1 - u) * y[i], t1 = u * u * (3 - 2 * u) * y[i + 1];
  const ld t2 = u * (1 - u) * (1 - u) * dx * s[i], t3 = -u * u * (1 - u) * dx * s[i + 1];
  *mag = fabsl(t0) + fabsl(t1) + fabsl(t2) + fabsl(t3);
  return t0 + t1 + t2 + t3;
}

int main() {
  int bad = 0;
  double worst = 0.;
  for (int nx : {2, 3, 5, 60, 100, 300, 2048}) {
    std::vector<double> x(nx), y(nx);
    for (int i = 0; i < nx; ++i) {
      x[i] = 1e-2 * pow(1e7, nx > 1 ? (double)i / (nx - 1) : 0.);        // s grid of TophatVariance on k in [1e-5, 1e2]
      y[i] = pow(x[i], -1.3) * (1. + 0.3 * sin(3. * log(x[i]))) / (1. + x[i] * x[i] * 1e-2);
    }
    for (int bc = 0; bc < 2; ++bc) {
      const std::vector<ld> s = full_slopes(x, y, bc);
      for (int W : {40, 64, 4096}) {
        if (W == 40 && nx != 2048) continue;   // a 40-knot window needs the fine grid (ordinates within ~3 % per knot)
        const int LW = 2 * W + 2 < nx ? 2 * W + 2 : nx;
        std::vector<double> w(LW), work(2 * LW);
        std::vector<double> qs = {x[0], x[nx - 1], 0.5 * (x[0] + x[1]), 0.5 * (x[nx - 2] + x[nx - 1])};
        for (int j = 0; j < 40; ++j) qs.push_back(x[0] * pow(x[nx - 1] / x[0], (j + 0.37) / 40.));
        for (int j = 1; j <= 20; ++j) qs.push_back((double)j);           // r = 1..20 Mpc/h
        for (double xv : qs) {
          if (xv < x[0] || xv > x[nx - 1]) continue;
          int first;
          const int L = spline_window_weights(x.data(), nx, bc, W, xv, w.data(), work.data(), &first);
          int skip;
          const int Lt = spline_trim_weights(w.data(), L, &skip);
          ld acc = 0, mag = 0;
          for (int j = skip; j < skip + Lt; ++j) { acc += (ld)w[j] * y[first + j]; mag += fabsl((ld)w[j] * y[first + j]); }
          ld hmag;
          const ld ref = full_eval(x, y, s, xv, &hmag);
          mag = fmaxl(mag, hmag);
          const double err = (double)(fabsl(acc - ref) / mag);
          if (err > worst) worst = err;
          if (!(err < 2e-14)) { ++bad; if (bad < 10) printf("nx %d bc %d W %d xv %g: %.17g vs %.17Lg err %.3e\n", nx, bc, W, xv, (double)acc, ref, err); }
        }
      }
    }
  }
  printf("worst relative error %.3e\n", worst);
  // the evaluation kernels' two-step form (spline_coeffs once per interval and column, spline_cubic_eval per point, with the interval's
  // reciprocal width from spline_query_kernel) against spline_poly: the same arithmetic, value and three derivatives
  {
    double worst2 = 0.;
    unsigned long long st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.; };
    for (int it = 0; it < 200000; ++it) {
      const double x0 = -3. + 6. * rnd(), x1 = x0 + 1e-3 + rnd(), y0 = -5. + 10. * rnd(), y1 = y0 + rnd() - 0.5, s0 = 4. * rnd() - 2., s1 = 4. * rnd() - 2.;
      const double xv = x0 + (x1 - x0) * (1.4 * rnd() - 0.2);
      const SplineCubic c = spline_coeffs(1. / (x1 - x0), y0, y1, s0, s1);
      for (int nu = 0; nu < 4; ++nu) {
        const double a = spline_poly(x0, x1, y0, y1, s0, s1, xv, nu), b = spline_cubic_eval(c, xv - x0, nu);
        const double err = fabs(a - b) / fmax(fabs(a), 1e-300);
        if (err > worst2) worst2 = err;
        if (!(a == b || err < 1e-15)) { ++bad; if (bad < 10) printf("two-step evaluation: nu %d %.17g vs %.17g\n", nu, a, b); }
      }
    }
    printf("two-step evaluation against spline_poly: worst relative difference %.3e\n", worst2);
  }
  if (bad) { printf("FAILED %d\n", bad); return 1; }
  printf("OK\n");
  return 0;
}
