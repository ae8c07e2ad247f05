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

// Row i of the slope system: lo * s_{i-1} + di * s_i + up * s_{i+1} = rhs_i.   bc: 0 natural, 1 clamped (zero slope),
// 2 not-a-knot (third derivative continuous at x_1 and x_{nx-2}; scipy/interpolate/_cubic.py, nx >= 4 required).
CPF_SHD void spline_row(const double* x, const int nx, const int bc, const int i, double& lo, double& di, double& up) {
  if (i == 0) {
    lo = 0.;
    if (bc == 1) { di = 1.; up = 0.; }
    else if (bc == 2) { di = x[2] - x[1]; up = x[2] - x[0]; }
    else { const double d0 = x[1] - x[0]; di = 2. * d0; up = d0; }
  } else if (i == nx - 1) {
    up = 0.;
    if (bc == 1) { di = 1.; lo = 0.; }
    else if (bc == 2) { di = x[nx - 2] - x[nx - 3]; lo = x[nx - 1] - x[nx - 3]; }
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

// One step of the factorisation shared by all columns: given c'_{i-1} (cprev) returns the four per-knot factors
//   Lw_i = lo_i / piv_i, cp_i = up_i / piv_i, P_i, Q_i with  rhs_i / piv_i = P_i (y_i - y_{i-1}) + Q_i (y_{i+1} - y_i)
// (not-a-knot end rows: P_0 (y_1 - y_0) + Q_0 (y_2 - y_1) and P_{n-1} (y_{n-2} - y_{n-3}) + Q_{n-1} (y_{n-1} - y_{n-2}))
// (differences of neighbouring ordinates are taken first, as scipy does, so smooth data keep their accuracy).
CPF_SHD void spline_factor_step(const double* x, const int nx, const int bc, const int i, double& cprev, double& Lw, double& P,
                                double& Q) {
  double lo, di, up;
  spline_row(x, nx, bc, i, lo, di, up);
  const double w = 1. / (di - lo * cprev);
  Lw = lo * w;
  cprev = up * w;
  if (bc == 2 && i == 0) {
    // rhs_0 = ((dx0 + 2 d) dx1 m_0 + dx0^2 m_1) / d, d = x_2 - x_0: applied to (y_1 - y_0) and (y_2 - y_1)
    const double dx0 = x[1] - x[0], dx1 = x[2] - x[1], d = x[2] - x[0];
    P = (dx0 + 2. * d) * dx1 / (d * dx0) * w;
    Q = dx0 * dx0 / (d * dx1) * w;
  } else if (bc == 2 && i == nx - 1) {
    // rhs_{n-1} = (dxl^2 m_{n-3} + (2 d + dxl) dxm m_{n-2}) / d, d = x_{n-1} - x_{n-3}: applied to (y_{n-2} - y_{n-3}) and (y_{n-1} - y_{n-2})
    const double dxl = x[nx - 1] - x[nx - 2], dxm = x[nx - 2] - x[nx - 3], d = x[nx - 1] - x[nx - 3];
    P = dxl * dxl / (d * dxm) * w;
    Q = (2. * d + dxl) * dxm / (d * dxl) * w;
  }
  else if (i == 0) { P = 0.; Q = bc == 1 ? 0. : 3. * w; }
  else if (i == nx - 1) { Q = 0.; P = bc == 1 ? 0. : 3. * w; }
  else {
    const double dm = x[i] - x[i - 1], dp = x[i + 1] - x[i];
    P = 3. * (dp / dm) * w;
    Q = 3. * (dm / dp) * w;
  }
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
  const double idx = 1. / (x1 - x0);      // one division per point (uniform over the columns), the rest are multiplications
  const double m = (y1 - y0) * idx;
  const double t = (s0 + s1 - 2. * m) * idx;
  const double c0 = t * idx, c1 = (m - s0) * idx - t, c2 = s0, c3 = y0;
  const double d = xv - x0;
  if (nu == 0) return c3 + d * (c2 + d * (c1 + d * c0));
  if (nu == 1) return c2 + d * (2. * c1 + d * 3. * c0);
  if (nu == 2) return 2. * c1 + 6. * c0 * d;
  return 6. * c0;
}


// The same polynomial in two steps, for kernels that evaluate many points per interval: the coefficients once per (interval, column) from
// idx = 1 / (x1 - x0) (a property of the interval: computed once per query point, not per column), then a Horner step per point at d = xv - x0.
struct SplineCubic { double c0, c1, c2, c3; };
CPF_SHD SplineCubic spline_coeffs(const double idx, const double y0, const double y1, const double s0, const double s1) {
  const double m = (y1 - y0) * idx;
  const double t = (s0 + s1 - 2. * m) * idx;
  SplineCubic c;
  c.c0 = t * idx; c.c1 = (m - s0) * idx - t; c.c2 = s0; c.c3 = y0;
  return c;
}
CPF_SHD double spline_cubic_eval(const SplineCubic& c, const double d, const int nu) {
  if (nu == 0) return c.c3 + d * (c.c2 + d * (c.c1 + d * c.c0));
  if (nu == 1) return c.c2 + d * (2. * c.c1 + d * 3. * c.c0);
  if (nu == 2) return 2. * c.c1 + 6. * c.c0 * d;
  return 6. * c.c0;
}

// ---- windowed evaluation weights -----------------------------------------------------------------------------------
// The spline value at xv is linear in the ordinates: S(xv) = sum_j w_j y_j.  The slope system is strictly diagonally
// dominant, so the influence of y_j on the slopes of the interval [x_i, x_{i+1}] holding xv decays like (2-sqrt 3)^|i-j|
// ~ 0.27^|i-j|: restricting the solve to the knots a = max(0, i-W) .. b = min(nx-1, i+1+W) changes S(xv) by O(0.27^W)
// (1e-23 for W = 40; W >= nx covers all knots and is the full solve).  The weights depend on the grid and on xv only,
// so they are computed once per query and shared by all rows.
//
// Split in three steps so that the kernel can run the last one (and the trimming) with a whole warp:
//   spline_window_setup  : window [a, b], Hermite factors of the bracketing interval
//   spline_window_solve  : z = T^-T (h10 e_p + h11 e_{p+1}) by Thomas elimination on the transposed window system (serial)
//   spline_window_weight : w_j = sum_m z_m d rhs_m / d y_j (+ h00, h01), independent for every j
// Truncated window ends use the natural end row (any consistent end row does: its effect is below the truncation).
struct SplineWindow {
  int a, L, p, bca, bcb;
  double h00, h01, h10, h11;
};

CPF_SHD SplineWindow spline_window_setup(const double* x, const int nx, const int bc, const int W, const double xv) {
  SplineWindow sw;
  const int i = spline_interval(x, nx, xv);
  sw.a = (i - W > 0) ? i - W : 0;
  const int b = (i + 1 + W < nx - 1) ? i + 1 + W : nx - 1;
  sw.L = b - sw.a + 1;
  sw.bca = (sw.a == 0) ? bc : 0;
  sw.bcb = (b == nx - 1) ? bc : 0;
  sw.p = i - sw.a;
  // Hermite form of the cubic on [x_i, x_{i+1}]: S = h00 y_i + h01 y_{i+1} + dx (h10 s_i + h11 s_{i+1})
  const double dx = x[i + 1] - x[i];
  const double u = (xv - x[i]) / dx;
  sw.h00 = (1. + 2. * u) * (1. - u) * (1. - u);
  sw.h01 = u * u * (3. - 2. * u);
  sw.h10 = u * (1. - u) * (1. - u) * dx;
  sw.h11 = -u * u * (1. - u) * dx;
  return sw;
}

CPF_SHD void spline_window_row(const double* xw, const SplineWindow& sw, const int m, double& lo, double& di, double& up) {
  if (m == 0) spline_row(xw, sw.L, sw.bca, 0, lo, di, up);
  else if (m == sw.L - 1) spline_row(xw, sw.L, sw.bcb, sw.L - 1, lo, di, up);
  else spline_row(xw, sw.L, 0, m, lo, di, up);
}

// cp, z: scratch / result of L doubles each.  Row m of T^T: up_{m-1} z_{m-1} + di_m z_m + lo_{m+1} z_{m+1} = rhs_m.
CPF_SHD void spline_window_solve(const double* xw, const SplineWindow& sw, double* cp, double* z) {
  const int L = sw.L;
  double lo_m, di_m, up_m, lo_n, di_n, up_n;   // rows m and m+1 of T
  double up_prev = 0., cprev = 0., zprev = 0.;
  spline_window_row(xw, sw, 0, lo_m, di_m, up_m);
  for (int m = 0; m < L; ++m) {
    if (m + 1 < L) spline_window_row(xw, sw, m + 1, lo_n, di_n, up_n); else { lo_n = 0.; di_n = 1.; up_n = 0.; }
    const double rhs = (m == sw.p ? sw.h10 : 0.) + (m == sw.p + 1 ? sw.h11 : 0.);
    const double piv = 1. / (di_m - up_prev * cprev);
    cprev = lo_n * piv;            // super-diagonal of T^T in row m is lo_{m+1}
    zprev = (rhs - up_prev * zprev) * piv;
    cp[m] = cprev;
    z[m] = zprev;
    up_prev = up_m;                // sub-diagonal of T^T in row m+1 is up_m
    lo_m = lo_n; di_m = di_n; up_m = up_n;
  }
  double zn = z[L - 1];
  for (int m = L - 2; m >= 0; --m) {
    zn = z[m] - cp[m] * zn;
    z[m] = zn;
  }
}

// rhs_m = 3 (dp (y_m - y_{m-1}) / dm + dm (y_{m+1} - y_m) / dp) inside, 3 (y_1 - y_0) / 3 (y_{L-1} - y_{L-2}) on natural
// end rows, 0 on clamped ones: collect the three rows that see y_j.
CPF_SHD double spline_window_weight(const double* xw, const SplineWindow& sw, const double* z, const int j) {
  const int L = sw.L;
  double w = 0.;
  if (j >= 1) {                                   // row m = j-1 sees y_j as y_{m+1}
    const int m = j - 1;
    if (m == 0) { if (sw.bca == 0) w += 3. * z[0]; }
    else { const double dm = xw[m] - xw[m - 1], dp = xw[m + 1] - xw[m]; w += 3. * z[m] * (dm / dp); }
  }
  if (j == 0) { if (sw.bca == 0) w -= 3. * z[0]; }
  else if (j == L - 1) { if (sw.bcb == 0) w += 3. * z[j]; }
  else { const double dm = xw[j] - xw[j - 1], dp = xw[j + 1] - xw[j]; w += 3. * z[j] * (dp / dm - dm / dp); }
  if (j + 1 <= L - 1) {                           // row m = j+1 sees y_j as y_{m-1}
    const int m = j + 1;
    if (m == L - 1) { if (sw.bcb == 0) w -= 3. * z[m]; }
    else { const double dm = xw[m] - xw[m - 1], dp = xw[m + 1] - xw[m]; w -= 3. * z[m] * (dp / dm); }
  }
  if (j == sw.p) w += sw.h00;
  if (j == sw.p + 1) w += sw.h01;
  return w;
}

// serial driver (CPU emulation): w[0..L-1] for the knots a..a+L-1, *first = a; `work` is scratch of 2*L doubles
CPF_SHD int spline_window_weights(const double* x, const int nx, const int bc, const int W, const double xv, double* w,
                                  double* work, int* first) {
  const SplineWindow sw = spline_window_setup(x, nx, bc, W, xv);
  spline_window_solve(x + sw.a, sw, work, work + sw.L);
  for (int j = 0; j < sw.L; ++j) w[j] = spline_window_weight(x + sw.a, sw, work + sw.L, j);
  *first = sw.a;
  return sw.L;
}

// Drop leading / trailing weights below 1e-40 of the largest one (they cannot reach the last bit of an fp64 sum unless
// the ordinates span more than 24 decades): returns the trimmed length, *skip = number of leading weights dropped.
CPF_SHD int spline_trim_weights(const double* w, const int L, int* skip) {
  double wmax = 0.;
  for (int j = 0; j < L; ++j) wmax = fmax(wmax, fabs(w[j]));
  const double thr = 1e-40 * wmax;
  int lo = 0, hi = L - 1;
  while (lo < hi && !(fabs(w[lo]) > thr)) ++lo;
  while (hi > lo && !(fabs(w[hi]) > thr)) --hi;
  *skip = lo;
  return hi - lo + 1;
}

}  // namespace cpf
