// cpf_spline_core.h — scalar building blocks of the cubic-spline kernels (host + device, so tests/emul can run them
// on the CPU).  See cpf_spline.cu for the formulation.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CPF_SHD __host__ __device__ __forceinline__
#else
#define CPF_SHD inline
#endif

namespace cpf {

// Row i of the slope system: lo * s_{i-1} + di * s_i + up * s_{i+1} = rhs_i.   bc: 0 natural, 1 clamped (zero slope).
CPF_SHD void spline_row(const double* x, const int nx, const int bc, const int i, double& lo, double& di, double& up) {
  if (i == 0) {
    lo = 0.;
    if (bc == 1) { di = 1.; up = 0.; }
    else { const double d0 = x[1] - x[0]; di = 2. * d0; up = d0; }
  } else if (i == nx - 1) {
    up = 0.;
    if (bc == 1) { di = 1.; lo = 0.; }
    else { const double dl = x[nx - 1] - x[nx - 2]; di = 2. * dl; lo = dl; }
  } else {
    const double dm = x[i] - x[i - 1], dp = x[i + 1] - x[i];
    lo = dp;
    di = 2. * (dm + dp);
    up = dm;
  }
}

// Right-hand side of row i from the three ordinates around knot i (ym = y_{i-1}, y0 = y_i, yp = y_{i+1}).
CPF_SHD double spline_rhs(const double* x, const int nx, const int bc, const int i, const double ym, const double y0, const double yp) {
  if (i == 0) return bc == 1 ? 0. : 3. * (yp - y0);
  if (i == nx - 1) return bc == 1 ? 0. : 3. * (y0 - ym);
  const double dm = x[i] - x[i - 1], dp = x[i + 1] - x[i];
  return 3. * (dp * ((y0 - ym) / dm) + dm * ((yp - y0) / dp));
}

// Interval index i with x_i <= xv < x_{i+1}, clamped to [0, nx-2] (so xv == x_{nx-1} and extrapolated points use the
// end polynomials, as scipy's PPoly does).
CPF_SHD int spline_interval(const double* x, const int nx, const double xv) {
  int lo = 0, hi = nx - 1;          // invariant: x[lo] <= xv < x[hi] when inside
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xv >= x[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

// nu-th derivative at xv of the cubic on [x0, x1] with end values y0, y1 and end slopes s0, s1, in scipy's power
// basis around x0:  c3 = y0, c2 = s0, c1 = (m - s0)/dx - t, c0 = t/dx, t = (s0 + s1 - 2 m)/dx, m = (y1 - y0)/dx.
CPF_SHD double spline_poly(const double x0, const double x1, const double y0, const double y1, const double s0, const double s1,
                           const double xv, const int nu) {
  const double dx = x1 - x0;
  const double m = (y1 - y0) / dx;
  const double t = (s0 + s1 - 2. * m) / dx;
  const double c0 = t / dx, c1 = (m - s0) / dx - t, c2 = s0, c3 = y0;
  const double d = xv - x0;
  if (nu == 0) return c3 + d * (c2 + d * (c1 + d * c0));
  if (nu == 1) return c2 + d * (2. * c1 + d * 3. * c0);
  if (nu == 2) return 2. * c1 + 6. * c0 * d;
  return 6. * c0;
}

}  // namespace cpf
