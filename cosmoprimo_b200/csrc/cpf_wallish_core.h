// cpf_wallish_core.h — per-thread phase functions of the fused Wallish2018 kernel (cosmoprimo/bao_filter.py:361-413):
// log(k P) -> DST-II -> even/odd second derivatives -> argmax boxes -> cut + re-spline -> DST-III -> exp(.)/k.
// One CTA of 256 threads processes TWO spectra (columns) at once, packed as the real and imaginary parts of one
// complex sequence, entirely in shared memory.  Host + device code: tests/emul/emul_wallish.cpp runs the same
// functions thread by thread on the CPU.
//
// Mathematics:
//  * DST-II of x (N = 4096) = DCT-II of x'_n = (-1)^n x_n read backwards; DCT-II by Makhoul's N-point FFT of
//    v[n] = x'[2n], v[N-1-n] = x'[2n+1]:  C_k = 2 Re(e^{-i pi k/2N} V_k).  Two columns share one complex FFT
//    (z = v_a + i v_b, V_a = (Z_k + conj Z_{N-k})/2, V_b = (Z_k - conj Z_{N-k})/2i).  Orthonormal scaling
//    sqrt(1/2N), last coefficient sqrt(1/4N) (scipy norm='ortho').  DST-III is the exact inverse of these steps.
//  * clamped cubic spline on the uniform knots 1..n: slopes solve s_{i-1} + 4 s_i + s_{i+1} = 3 (y_{i+1} - y_{i-1}),
//    s_0 = s_{n-1} = 0.  The Thomas pivots depend on i only (and converge to 2 - sqrt 3 within 20 rows); influence of a
//    right-hand side decays by 0.27 per knot, so each thread eliminates its own 16-knot chunk after a 32-knot warm-up
//    (error 5e-19): fully parallel, no scratch beyond one array.
//  * cut + re-spline: the second spline (bao_filter.py:400-402) differs from the first only by the removed box, so
//    only the two slopes at the knots bounding the box are needed: two short one-sided eliminations + a 2x2 solve.
#pragma once

#include "cpf_fft_core.h"

namespace cpf {

struct WallishGeo {
  static constexpr int N = 4096;            // DST length (bao_filter.py:364)
  static constexpr int H = 2048;            // even / odd sequence length
  static constexpr int T = 256;             // threads per CTA
  static constexpr int CH = 16;             // knots per thread in the chunked eliminations
  static constexpr int WARM = 32;           // warm-up knots: (2 - sqrt 3)^32 = 5e-19, below the rounding of the exact solve
  static constexpr int HP = H + H / CH + 4;   // padded half length (one pad element per 16: conflict-free chunk access; + 4: the even and the odd
                                              // sequence sit 4 bank groups apart, accesses that alternate between them are conflict-free too)
  static constexpr int BUF = 2 * HP;        // elements of one shared-memory array (>= 16*257 exchange elements)
  static constexpr int MARGIN_FIRST = 20, MARGIN_SECOND = 5, OFF_LO = -10, OFF_HI = 20;   // bao_filter.py:387-389
};

#define CPF_LAMBDA 0.26794919243112270647   // 2 - sqrt(3): limit of the Thomas pivots 1/(4 - w)
// scipy norm='ortho' scaling of the DST-II of length N = 4096: sqrt(1/2N), last coefficient sqrt(1/4N)
#define CPF_DST_S 0.011048543456039806
#define CPF_DST_S_LAST 0.0078125

// padded position of knot i of parity h
CPF_HD int wpos(const int h, const int i) { return h * WallishGeo::HP + i + (i >> 4); }

// Thomas pivot w_i = 1/(d_i - l_i c_{i-1}) of the clamped uniform system, rows 0..n-1 (row 0 and n-1: s = 0)
CPF_HD double wpivot(const double* wtab, const int i, const int n) {
  if (i == 0 || i == n - 1) return 1.;
  return i < 32 ? wtab[i] : CPF_LAMBDA;
}
// c_i = u_i w_i with u_0 = u_{n-1} = 0
CPF_HD double wcp(const double* wtab, const int i, const int n) {
  if (i == 0 || i == n - 1) return 0.;
  return i < 32 ? wtab[i] : CPF_LAMBDA;
}

// ---- DST-II post-processing: Z (natural order, in A) -> orthonormal DST-II coefficients, de-interleaved + padded ----
// thread t owns bins k = t + 256 r (zk[r] in registers); tw[k] = exp(-i pi k / 2N)
CPF_HD void wallish_dst2_post(const int t, const double2 (&zk)[16], const double2* A, double2* X, const double2* tw) {
  typedef WallishGeo G;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k = t + G::T * r;
    const double2 z = zk[r], zm = A[(G::N - k) & (G::N - 1)];
    // V_a = (z + conj zm)/2, V_b = (z - conj zm)/(2i)
    const double2 va = mk2(0.5 * (z.x + zm.x), 0.5 * (z.y - zm.y));
    const double2 vb = mk2(0.5 * (z.y + zm.y), 0.5 * (zm.x - z.x));
    const double2 w = CPF_LDG(tw + k);
    const int kk = G::N - 1 - k;                       // DST-II index
    const double sc = 2. * (kk == G::N - 1 ? CPF_DST_S_LAST : CPF_DST_S);
    const double ca = sc * (w.x * va.x - w.y * va.y), cb = sc * (w.x * vb.x - w.y * vb.y);
    X[wpos(kk & 1, kk >> 1)] = mk2(ca, cb);
  }
}

// ---- DST-II post-processing with ONE buffer: the thread's bins zk[r] (k = t + 256 r) are replaced in registers by the DST-II
// coefficient of index N-1-k (A holds the bins in natural order); after a barrier wallish_dst2_store scatters them into the
// same buffer in the de-interleaved, padded layout the spline phases read.
CPF_HD void wallish_dst2_coef(const int t, double2 (&zk)[16], const double2* A, const double2* tw) {
  typedef WallishGeo G;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k = t + G::T * r;
    const double2 z = zk[r], zm = A[(G::N - k) & (G::N - 1)];
    const double2 va = mk2(0.5 * (z.x + zm.x), 0.5 * (z.y - zm.y));
    const double2 vb = mk2(0.5 * (z.y + zm.y), 0.5 * (zm.x - z.x));
    const double2 w = CPF_LDG(tw + k);
    const double sc = 2. * (k == 0 ? CPF_DST_S_LAST : CPF_DST_S);
    zk[r] = mk2(sc * (w.x * va.x - w.y * va.y), sc * (w.x * vb.x - w.y * vb.y));
  }
}

CPF_HD void wallish_dst2_store(const int t, const double2 (&zk)[16], double2* X) {
  typedef WallishGeo G;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int kk = G::N - 1 - (t + G::T * r);
    X[wpos(kk & 1, kk >> 1)] = zk[r];
  }
}

// ---- clamped splines through the even / odd coefficients, chunks in registers ----------------------------------------
// Y: knot values (padded layout, parity h).  sq != 0: the values are y_i * x_i^2.
CPF_HD double2 wallish_y(const double2* Y, const int h, const int i, const int sq) {
  double2 y = Y[wpos(h, i)];
  if (sq) { const double x2 = (double)(i + 1) * (double)(i + 1); y.x *= x2; y.y *= x2; }
  return y;
}

// Thread t = 128 h + c owns the 16 knots first = 16 c .. first + 15 of parity h and keeps their reduced right-hand sides /
// slopes / second derivatives in d[16] (registers; no second and third shared-memory array: two CTAs fit one SM).
// The recurrence d_i = (r_i - lo_i d_{i-1}) w_i is affine in the value flowing in from the left chunk:
//   d_i = d_i(0) + m_i d_in,  m_i = prod_{j <= i} (-lo_j w_j),  m_15 ~ (-0.27)^16 = 7e-10.
// Step 1 (wallish_forward_local): every thread runs its chunk from zero inflow and publishes (E, M) = (d_15(0), m_15).
// Step 2 (wallish_forward_fix): the true inflow is E_{c-1} + M_{c-1} E_{c-2} (the next term, ~ 5e-19 E, is below the rounding of the
// exact solve); the chunk is corrected by d_i += m_i d_in instead of being run again.  The back substitution
// s_i = d_i - cp_i s_{i+1} is treated the same way from the right.
CPF_HD void wallish_forward_local(const int t, const double2* X, double2 (&d)[16], double2* E, double* M, const double* wtab) {
  typedef WallishGeo G;
  const int h = t >> 7, first = (t & 127) * G::CH;
  double m = 1.;
  double2 dp = mk2(0., 0.);
  double2 ym = first > 0 ? X[wpos(h, first - 1)] : mk2(0., 0.), y0 = X[wpos(h, first)];
#pragma unroll
  for (int j = 0; j < G::CH; ++j) {
    const int i = first + j;
    const double2 yp = i + 1 < G::H ? X[wpos(h, i + 1)] : mk2(0., 0.);
    const bool edge = (i == 0 || i == G::H - 1);
    const double w = wpivot(wtab, i, G::H);
    const double rx = edge ? 0. : 3. * (yp.x - ym.x), ry = edge ? 0. : 3. * (yp.y - ym.y);
    const double lo = edge ? 0. : 1.;
    dp = mk2((rx - lo * dp.x) * w, (ry - lo * dp.y) * w);
    m *= -lo * w;
    d[j] = dp;
    ym = y0; y0 = yp;
  }
  E[t] = dp;
  M[t] = m;
}

// SYNC hook of the two fix-and-backward steps: called between the last read of (E, M) and the write of (Eb, Mb).  The kernel passes a CTA
// barrier and lets Eb / Mb share the memory of E / M (6 KB less shared memory: three CTAs per SM); the CPU emulation keeps separate arrays.
struct WallishNoSync { CPF_HD void operator()() const {} };

// inflow correction of the forward pass, then the back substitution from zero inflow: d[j] <- s_j(0); publishes (Eb, Mb)
template <class SYNC = WallishNoSync>
CPF_HD void wallish_forward_fix_backward_local(const int t, double2 (&d)[16], const double2* E, const double* M, double2* Eb, double* Mb,
                                               const double* wtab, const SYNC& sync = SYNC()) {
  typedef WallishGeo G;
  const int c = t & 127, first = c * G::CH;
  double2 din = mk2(0., 0.);
  if (c >= 1) {
    din = E[t - 1];
    if (c >= 2) { din.x += M[t - 1] * E[t - 2].x; din.y += M[t - 1] * E[t - 2].y; }
  }
  double m = 1.;
#pragma unroll
  for (int j = 0; j < G::CH; ++j) {
    const int i = first + j;
    const bool edge = (i == 0 || i == G::H - 1);
    m *= edge ? 0. : -wpivot(wtab, i, G::H);
    d[j].x = fma(m, din.x, d[j].x);
    d[j].y = fma(m, din.y, d[j].y);
  }
  double2 s = mk2(0., 0.);
  m = 1.;
#pragma unroll
  for (int j = G::CH - 1; j >= 0; --j) {
    const double cp = wcp(wtab, first + j, G::H);
    s = mk2(d[j].x - cp * s.x, d[j].y - cp * s.y);
    m *= -cp;
    d[j] = s;
  }
  sync();
  Eb[t] = s;
  Mb[t] = m;
}

// best second derivative of a thread's 16-knot chunk inside the search range [MARGIN_FIRST, H - MARGIN_FIRST) of both
// columns (index -1: no knot of the chunk is in range); ties keep the lowest index (numpy argmax)
struct WallishBest {
  double vx, vy;
  int ix, iy;
};

CPF_HD void wallish_best_update(WallishBest& b, const double x, const double y, const int i, const int lox, const int loy) {
  const int hi = WallishGeo::H - WallishGeo::MARGIN_FIRST;
  if (i >= lox && i < hi && (b.ix < 0 || x > b.vx || (x == b.vx && i < b.ix))) { b.vx = x; b.ix = i; }
  if (i >= loy && i < hi && (b.iy < 0 || y > b.vy || (y == b.vy && i < b.iy))) { b.vy = y; b.iy = i; }
}

// inflow correction of the back substitution fused with the second derivatives at the knots (bao_filter.py:379, 382):
// d[j] <- dd_{first + j}; returns the chunk's best
CPF_HD WallishBest wallish_backward_dd(const int t, const double2* X, double2 (&d)[16], const double2* Eb, const double* Mb, const double* wtab) {
  typedef WallishGeo G;
  const int h = t >> 7, c = t & 127;
  const int first = c * G::CH, last = first + G::CH - 1;
  double2 sin_ = mk2(0., 0.);           // s_{last+1}: slope flowing in from the right (nothing beyond the last knot)
  if (c <= 126) {
    sin_ = Eb[t + 1];
    if (c <= 125) { sin_.x += Mb[t + 1] * Eb[t + 2].x; sin_.y += Mb[t + 1] * Eb[t + 2].y; }
  }
  double2 sn = sin_;
  double2 yn = last + 1 < G::H ? X[wpos(h, last + 1)] : mk2(0., 0.);
  double m = 1.;
  double2 ddlast = mk2(0., 0.);
  WallishBest b;
  b.vx = b.vy = 0.; b.ix = b.iy = -1;
#pragma unroll
  for (int j = G::CH - 1; j >= 0; --j) {
    const int i = first + j;
    m *= -wcp(wtab, i, G::H);
    const double2 yi = X[wpos(h, i)];
    const double2 s = mk2(fma(m, sin_.x, d[j].x), fma(m, sin_.y, d[j].y));
    if (i < G::H - 1) {
      // dd_i = 2 c1 = 2 (3 m_i - 2 s_i - s_{i+1}), m_i = y_{i+1} - y_i; last knot: -6 m_{n-2} + 2 s_{n-2} + 4 s_{n-1}
      const double mx = yn.x - yi.x, my = yn.y - yi.y;
      const double2 dd = mk2(2. * (3. * mx - 2. * s.x - sn.x), 2. * (3. * my - 2. * s.y - sn.y));
      d[j] = dd;
      wallish_best_update(b, dd.x, dd.y, i, G::MARGIN_FIRST, G::MARGIN_FIRST);
      if (i == G::H - 2) ddlast = mk2(-6. * mx + 2. * s.x + 4. * sn.x, -6. * my + 2. * s.y + 4. * sn.y);
    }
    sn = s;
    yn = yi;
  }
  if (last == G::H - 1) d[G::CH - 1] = ddlast;
  return b;
}

// candidate of a thread's chunk for the second search range [lb, H - MARGIN_FIRST) (lb per column): the chunk best when
// the whole chunk lies above lb, a re-scan of the chunk's knots >= lb when lb falls inside it, nothing below
CPF_HD WallishBest wallish_chunk_candidate(const int t, const double2 (&dd)[16], const int lbx, const int lby, const WallishBest& chunk) {
  typedef WallishGeo G;
  const int first = (t & 127) * G::CH, last = first + G::CH - 1;
  WallishBest b;
  b.vx = b.vy = 0.; b.ix = b.iy = -1;
  if (first >= lbx) { b.vx = chunk.vx; b.ix = chunk.ix; }
  if (first >= lby) { b.vy = chunk.vy; b.iy = chunk.iy; }
  if ((first < lbx && last >= lbx) || (first < lby && last >= lby)) {
    const int lox = first < lbx ? lbx : G::H, loy = first < lby ? lby : G::H;     // columns already settled are skipped
#pragma unroll
    for (int j = 0; j < G::CH; ++j) wallish_best_update(b, dd[j].x, dd[j].y, first + j, lox, loy);
  }
  return b;
}

// merge rule of the reductions: keep the larger value, the lower index on ties, ignore empty candidates
CPF_HD void wallish_best_merge(double& v, int& i, const double ov, const int oi) {
  if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
}

// ---- cut + re-spline: slopes at the knots L = b0-1 and R = b1+1 bounding the removed box (values y x^2) ------------
struct WallishGap {
  int b0, b1;          // removed index range [b0, b1]
  double sL, sR;       // slopes at L and R
  double yL, m, G;     // value at L, chord slope over the gap, gap width
  int ok;              // 0: box reaches the end of the array (reference yields NaN there)
};

// The two one-sided eliminations of a sequence run over at most WARM rows each: left over rows < L ascending (unaffected by
// the cut), right over rows > R descending (mirrored system: the pivots are those of row n-1-i).  Split into pieces so that
// the kernel computes the right-hand sides with all threads, runs the 8 short chains (4 sequences x 2 sides) on 8 threads
// and finishes with the 2x2 solves; wallish_gap_solve chains the same pieces serially (CPU emulation, reference flow).
CPF_HD int wallish_gap_row(const int b0, const int b1, const int side, const int step) {
  typedef WallishGeo Gm;
  const int n = Gm::H, L = b0 - 1, R = b1 + 1;
  if (side == 0) {
    const int i = (L - Gm::WARM > 0 ? L - Gm::WARM : 0) + step;
    return i < L ? i : -1;
  }
  const int i = (R + Gm::WARM < n - 1 ? R + Gm::WARM : n - 1) - step;
  return i > R ? i : -1;
}

CPF_HD double wallish_gap_y(const double2* X, const int h, const int col, const int i) {
  const double2 y = wallish_y(X, h, i, 1);
  return col ? y.y : y.x;
}

CPF_HD bool wallish_gap_ok(const int b0, const int b1) { return b0 - 1 >= 1 && b1 + 1 <= WallishGeo::H - 2; }

// right-hand side of row i of the y x^2 system (0 on the clamped end rows and for padding steps, i < 0)
CPF_HD double wallish_gap_rhs(const double2* X, const int h, const int col, const int i) {
  if (i <= 0 || i >= WallishGeo::H - 1) return 0.;
  return 3. * (wallish_gap_y(X, h, col, i + 1) - wallish_gap_y(X, h, col, i - 1));
}

// reduced right-hand side at the last eliminated row; r[step] = wallish_gap_rhs of row wallish_gap_row(.., step)
CPF_HD double wallish_gap_chain(const double* r, const int b0, const int b1, const int side, const double* wtab) {
  typedef WallishGeo Gm;
  const int n = Gm::H;
  double d = 0.;
  for (int step = 0; step < Gm::WARM; ++step) {
    const int i = wallish_gap_row(b0, b1, side, step);
    if (i < 0) break;
    const bool edge = (i == 0 || i == n - 1);
    d = (r[step] - (edge ? 0. : 1.) * d) * wpivot(wtab, side ? n - 1 - i : i, n);
  }
  return d;
}

CPF_HD WallishGap wallish_gap_finish(const double2* X, const int h, const int col, const int b0, const int b1, const double dpL,
                                     const double dqR, const double* wtab) {
  typedef WallishGeo Gm;
  WallishGap g;
  g.b0 = b0; g.b1 = b1;
  const int n = Gm::H, L = b0 - 1, R = b1 + 1;
  g.ok = wallish_gap_ok(b0, b1) ? 1 : 0;
  g.sL = g.sR = g.yL = g.m = 0.; g.G = 1.;
  if (!g.ok) return g;
  const double cpL = wcp(wtab, L - 1, n), cqR = wcp(wtab, n - 1 - (R + 1), n);
  const double G = (double)(R - L);
  const double yL = wallish_gap_y(X, h, col, L), yR = wallish_gap_y(X, h, col, R);
  const double mleft = yL - wallish_gap_y(X, h, col, L - 1), mgap = (yR - yL) / G, mright = wallish_gap_y(X, h, col, R + 1) - yR;
  const double rhsL = 3. * (G * mleft + mgap), rhsR = 3. * (mgap + G * mright);
  const double a11 = 2. * (1. + G) - G * cpL, a22 = 2. * (G + 1.) - G * cqR;
  const double b1_ = rhsL - G * dpL, b2_ = rhsR - G * dqR;
  const double det = a11 * a22 - 1.;
  g.sL = (b1_ * a22 - b2_) / det;
  g.sR = (a11 * b2_ - b1_) / det;
  g.yL = yL; g.m = mgap; g.G = G;
  return g;
}

CPF_HD WallishGap wallish_gap_solve(const double2* X, const int h, const int col, const int b0, const int b1, const double* wtab) {
  typedef WallishGeo Gm;
  double d[2] = {0., 0.};
  if (wallish_gap_ok(b0, b1)) {
    double r[Gm::WARM];
    for (int side = 0; side < 2; ++side) {
      for (int step = 0; step < Gm::WARM; ++step) r[step] = wallish_gap_rhs(X, h, col, wallish_gap_row(b0, b1, side, step));
      d[side] = wallish_gap_chain(r, b0, b1, side, wtab);
    }
  }
  return wallish_gap_finish(X, h, col, b0, b1, d[0], d[1], wtab);
}

// new value of knot i of a sequence after cut + re-spline: spline(x_i)/x_i^2 (bao_filter.py:402)
CPF_HD double wallish_fill(const double y, const int i, const WallishGap& g) {
  const double x = (double)(i + 1), x2 = x * x;
  if (!g.ok) {
    // box reaches the array end: the reference's spline is not defined beyond its last knot (extrapolate=False)
    return i >= g.b0 ? nan("") : y;
  }
  if (i < g.b0 || i > g.b1) return y;   // the reference's (y x^2) / x^2 differs from y by at most one ulp
  const double t = (g.sL + g.sR - 2. * g.m) / g.G;
  const double c0 = t / g.G, c1 = (g.m - g.sL) / g.G - t;
  const double d = (double)(i - (g.b0 - 1));
  return (g.yL + d * (g.sL + d * (c1 + d * c0))) / x2;
}

// ---- DST-III pre-processing: orthonormal coefficients X -> FFT input Z for bins k = t + 256 r -----------------------
CPF_HD void wallish_dst3_pre(const int t, const double2* X, double2 (&zk)[16], const double2* tw) {
  typedef WallishGeo G;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k = t + G::T * r;
    // C'_k = X[N-1-k]/s, C'_{N-k} = X[k-1]/s (0 for k = 0)
    const int k1 = G::N - 1 - k, k2 = k - 1;
    const double f1 = 0.5 / (k1 == G::N - 1 ? CPF_DST_S_LAST : CPF_DST_S);     // the 0.5 of V_k folded in
    double2 c1 = X[wpos(k1 & 1, k1 >> 1)];
    c1.x *= f1; c1.y *= f1;
    double2 c2 = mk2(0., 0.);
    if (k2 >= 0) {
      const double f2 = 0.5 / (k2 == G::N - 1 ? CPF_DST_S_LAST : CPF_DST_S);
      c2 = X[wpos(k2 & 1, k2 >> 1)];
      c2.x *= f2; c2.y *= f2;
    }
    const double2 w = CPF_LDG(tw + k);                 // exp(-i pi k/2N); need 0.5 * conj(w) * (C'_k - i C'_{N-k})
    // column a: (c1.x - i c2.x) * conj(w) ; column b: (c1.y - i c2.y) * conj(w)   [0.5 folded into c1, c2]
    const double var = w.x * c1.x - w.y * c2.x, vai = -w.x * c2.x - w.y * c1.x;   // conj(w) = (w.x, -w.y)
    const double vbr = w.x * c1.y - w.y * c2.y, vbi = -w.x * c2.y - w.y * c1.y;
    zk[r] = mk2(var - vbi, vai + vbr);                  // Z = V_a + i V_b
  }
}

// output index of FFT bin m of the DST-III transform and its sign (x = (-1)^j x'_j, x'[2n] = v[n], x'[2n+1] = v[N-1-n],
// v[n] = F[(N-n) mod N] / N)
CPF_HD int wallish_dst3_out_index(const int m, double& sign) {
  typedef WallishGeo G;
  if (m == 0) { sign = 1.; return 0; }
  if (m <= G::N / 2) { sign = -1.; return 2 * m - 1; }
  sign = 1.;
  return 2 * (G::N - m);
}

// ---- final stage (bao_filter.py:415-423) inside the fused kernel ----------------------------------------------------
// The spliced clamped spline -- nl unfiltered knots (self.k < 5e-4), the filtered spectrum on 1e-2 < k < 1.5, nr unfiltered knots
// (self.k > 2) -- is solved per pair of spectra in shared memory and evaluated at self.k.  Only the lz <= 64 / rz <= 64 edge knots next
// to the filtered ones are kept: the slope system is strictly diagonally dominant, the influence of a knot decays by >= 2 per knot
// (2^-64 = 5e-20), and the output wavenumbers outside [5e-4, 2] are knots themselves (the spline returns pk there, wiggles = 1).
// Knot c of the nc kept ones sits at ypos(c) (one pad element per 16: conflict-free chunk access); thread t owns the knots
// 16 t .. 16 t + 15.  The per-knot factors Lw, cp, P, Q of cpf_spline_core.h (spline_factor_step) depend on the knots only and come
// from a table in thread-major order: factor f of knot 16 t + j at facT[(4 j + f) * 256 + t] (coalesced; zeros beyond nc).
// Same two-step scheme as above with four inflow terms (non-uniform knots: |Lw|, |cp| <= 1/2, M <= 1.5e-5 per chunk, M^4 = 5e-20).
CPF_HD int ypos(const int i) { return i + (i >> 4); }

CPF_HD double2 wallish_inflow4(const double2* E, const double* M, const int t, const int dir) {
  // dir = -1: E_{t-1} + M_{t-1} (E_{t-2} + M_{t-2} (E_{t-3} + M_{t-3} E_{t-4})), chunks below 0 / above 255 do not exist
  double2 acc = mk2(0., 0.);
#pragma unroll
  for (int j = 4; j >= 1; --j) {
    const int c = t + dir * j;
    if (c < 0 || c > WallishGeo::T - 1) continue;
    const double m = j < 4 ? M[c] : 0.;
    const double2 e = E[c];
    acc = mk2(fma(m, acc.x, e.x), fma(m, acc.y, e.y));
  }
  return acc;
}

// Factors of the final solve.  Where the knots are uniform (the filtered ones, away from the splice points) the factors have converged to
// constants: threads t0 <= t < t1 take them from here instead of loading 48 table entries each (the table was 130 KB of L2 -> SM
// traffic per pair of spectra, three times the spectra themselves).
struct WallishFinFac {
  const double* facT;
  int t0, t1;
  double Lw, cp, P, Q;
};

template <bool UNI>
CPF_HD void wallish_fin_forward_local_t(const int t, const int nc, const double2* Y, const WallishFinFac& fc, double2 (&d)[16], double2* E, double* M) {
  typedef WallishGeo G;
  const int first = t * G::CH;
  double m = 1.;
  double2 dp = mk2(0., 0.);
  double2 ym = (first > 0 && first - 1 < nc) ? Y[ypos(first - 1)] : mk2(0., 0.), y0 = first < nc ? Y[ypos(first)] : mk2(0., 0.);
#pragma unroll
  for (int j = 0; j < G::CH; ++j) {
    const int i = first + j;
    const double2 yp = i + 1 < nc ? Y[ypos(i + 1)] : mk2(0., 0.);
    const double Lw = UNI ? fc.Lw : CPF_LDG(fc.facT + (4 * j + 0) * G::T + t);
    const double P = UNI ? fc.P : CPF_LDG(fc.facT + (4 * j + 2) * G::T + t), Q = UNI ? fc.Q : CPF_LDG(fc.facT + (4 * j + 3) * G::T + t);
    const double ax = y0.x - ym.x, ay = y0.y - ym.y, bx = yp.x - y0.x, by = yp.y - y0.y;
    dp = mk2(fma(P, ax, fma(Q, bx, -Lw * dp.x)), fma(P, ay, fma(Q, by, -Lw * dp.y)));
    m *= -Lw;
    d[j] = dp;
    ym = y0; y0 = yp;
  }
  E[t] = dp;
  M[t] = m;
}

CPF_HD void wallish_fin_forward_local(const int t, const int nc, const double2* Y, const WallishFinFac& fc, double2 (&d)[16], double2* E, double* M) {
  if (t >= fc.t0 && t < fc.t1) wallish_fin_forward_local_t<true>(t, nc, Y, fc, d, E, M);
  else wallish_fin_forward_local_t<false>(t, nc, Y, fc, d, E, M);
}

template <bool UNI>
CPF_HD void wallish_fin_fix_backward_local_t(const int t, const WallishFinFac& fc, double2 (&d)[16], const double2* E, const double* M, double2& s_out, double& m_out) {
  typedef WallishGeo G;
  const double2 din = wallish_inflow4(E, M, t, -1);
  double m = 1.;
#pragma unroll
  for (int j = 0; j < G::CH; ++j) {
    m *= -(UNI ? fc.Lw : CPF_LDG(fc.facT + (4 * j + 0) * G::T + t));
    d[j].x = fma(m, din.x, d[j].x);
    d[j].y = fma(m, din.y, d[j].y);
  }
  double2 s = mk2(0., 0.);
  m = 1.;
#pragma unroll
  for (int j = G::CH - 1; j >= 0; --j) {
    const double cp = UNI ? fc.cp : CPF_LDG(fc.facT + (4 * j + 1) * G::T + t);
    s = mk2(fma(-cp, s.x, d[j].x), fma(-cp, s.y, d[j].y));
    m *= -cp;
    d[j] = s;
  }
  s_out = s;
  m_out = m;
}

template <class SYNC = WallishNoSync>
CPF_HD void wallish_fin_fix_backward_local(const int t, const WallishFinFac& fc, double2 (&d)[16], const double2* E, const double* M, double2* Eb, double* Mb,
                                           const SYNC& sync = SYNC()) {
  double2 s;
  double m;
  if (t >= fc.t0 && t < fc.t1) wallish_fin_fix_backward_local_t<true>(t, fc, d, E, M, s, m);
  else wallish_fin_fix_backward_local_t<false>(t, fc, d, E, M, s, m);
  sync();
  Eb[t] = s;
  Mb[t] = m;
}

// d[j] <- slope at knot 16 t + j
template <bool UNI>
CPF_HD void wallish_fin_backward_fix_t(const int t, const WallishFinFac& fc, double2 (&d)[16], const double2* Eb, const double* Mb) {
  typedef WallishGeo G;
  const double2 sin_ = wallish_inflow4(Eb, Mb, t, +1);
  double m = 1.;
#pragma unroll
  for (int j = G::CH - 1; j >= 0; --j) {
    m *= -(UNI ? fc.cp : CPF_LDG(fc.facT + (4 * j + 1) * G::T + t));
    d[j].x = fma(m, sin_.x, d[j].x);
    d[j].y = fma(m, sin_.y, d[j].y);
  }
}

CPF_HD void wallish_fin_backward_fix(const int t, const WallishFinFac& fc, double2 (&d)[16], const double2* Eb, const double* Mb) {
  if (t >= fc.t0 && t < fc.t1) wallish_fin_backward_fix_t<true>(t, fc, d, Eb, Mb);
  else wallish_fin_backward_fix_t<false>(t, fc, d, Eb, Mb);
}

// slopes needed by the queries of round r go to their slots (slotT[(16 r + j) * 256 + t], -1: not needed)
CPF_HD void wallish_fin_scatter(const int t, const int round, const int* slotT, const double2 (&d)[16], double2* SL) {
  typedef WallishGeo G;
#pragma unroll
  for (int j = 0; j < G::CH; ++j) {
    const int sl = CPF_LDG(slotT + (G::CH * round + j) * G::T + t);
    if (sl >= 0) SL[sl] = d[j];
  }
}

// query q: qinfo[4 q ..] = {pos0, pos1, slot0, slot1} (pos0 = -1: self.k[q] is a spliced-in knot, the spline returns pk; -2: outside the
// knots, extrapolate=False gives NaN), qh[4 q ..] = Hermite factors h00, h01, h10 dx, h11 dx; th = Gaussian top-hat (:425-431)
CPF_HD double2 wallish_fin_eval(const int q, const int* qinfo, const double* qh, const double2* Y, const double2* SL, const double2 pk,
                                const double th) {
  const int pos0 = CPF_LDG(qinfo + 4 * q);
  if (pos0 == -1) return pk;
  if (pos0 < 0) return mk2(nan(""), nan(""));
  const int pos1 = CPF_LDG(qinfo + 4 * q + 1), sl0 = CPF_LDG(qinfo + 4 * q + 2), sl1 = CPF_LDG(qinfo + 4 * q + 3);
  const double h0 = CPF_LDG(qh + 4 * q), h1 = CPF_LDG(qh + 4 * q + 1), h2 = CPF_LDG(qh + 4 * q + 2), h3 = CPF_LDG(qh + 4 * q + 3);
  const double2 y0 = Y[pos0], y1 = Y[pos1], s0 = SL[sl0], s1 = SL[sl1];
  const double sa = fma(h0, y0.x, fma(h1, y1.x, fma(h2, s0.x, h3 * s1.x)));         // :420
  const double sb = fma(h0, y0.y, fma(h1, y1.y, fma(h2, s0.y, h3 * s1.y)));
  return mk2(pk.x / ((pk.x / sa - 1.) * th + 1.), pk.y / ((pk.y / sb - 1.) * th + 1.));     // :422-423
}

}  // namespace cpf
