// cpf_stream_core.h — per-thread phases of the "stream" FFTLog kernels (N = 4096, 256 (virtual) threads per pair of
// rows), as __host__ __device__ straight-line code so that tests/emul/emul_stream.cpp can run them thread by thread
// on the CPU.
//
// Data flow of one pair of rows (thread tau = 16 H + L, 16 complex values per thread in registers).
// The two forward FFTs of cpf_fftlog.cu are factored 16 x 16 x 16 with these thread roles:
//
//   FFT #1   P1  thread (m1,m2) = (H,L): DFT over n1 of x[256 n1 + 16 m1 + m2], twiddle w_4096^{tau k1}   -> A[k1]
//            P2  thread (k1,m2) = (H,L): DFT over m1, twiddle w_256^{L l1}                                 -> B[l1]
//            P3  thread (k1,l1) = (H,L): DFT over m2                          -> X[H + 16 L + 256 l2] in register l2
//   kernel multiply X[.] *= ut[.]
//   FFT #2   P1' thread (m2',m1') = (H,L): DFT over l2, twiddle w_4096^{(H + 16 L) k1'}                   -> A'[k1']
//            P2' thread (m2',k1') = (H,L): DFT over m1', twiddle w_256^{H l1'}                             -> B'[l1']
//            P3' thread (l1',k1') = (H,L): DFT over m2'               -> g[tau + 256 l2'] in register l2' (natural order)
//
// Exchange buffer: slot(a,b,c) = a*RS + 17 b + c (complex elements), a,b,c in 0..15.  Every access below is
// bank-conflict-free for 128-bit accesses (a quarter-warp touches 8 consecutive c, or 8 consecutive b with 17 b = b mod 8).
//
//   phase   reads                      writes                     synchronisation before the reads
//   P1      -                          slot(k1, H, L)   all k1    -
//   P2      slot(H, m1, L)  all m1     slot(H, l1, L)   in place  group barrier (P1 writes come from every warp)
//   P3      slot(H, L, m2)  all m2     -                          warp  (row H belongs to one half-warp)
//   P1'     -                          slot(H, L, k1')  own slots -
//   P2'     slot(H, m1', L) all m1'    slot(H, l1', L)  in place  warp
//   P3'     slot(m2', H, L) all m2'    -                          group barrier
//   next P1 writes slot(k1, H, L) = exactly the slots this thread read in P3'.
//
// A thread only ever writes slots that it was itself the last reader of, so there is no write-after-read hazard and
// a pair of rows costs two group barriers and two warp barriers in total.
//
// Each phase is split into load / compute / store so that a kernel can interleave two independent column sets in
// one thread (fftlog_stream2_kernel); the one-set kernel and the emulation call the st_p* wrappers.
#pragma once

#include "cpf_fft_core.h"

namespace cpf {

constexpr int ST_RS = 273;                 // exchange-buffer row stride (complex): >= 17*15 + 15 + 1
constexpr int ST_GROUP_ELEMS = 16 * ST_RS;
CPF_HDC int st_slot(const int a, const int b, const int c) { return a * ST_RS + 17 * b + c; }

// ---- scaled twiddles (round 2) -------------------------------------------------------------------------------------
// An inter-pass twiddle w = exp(-i theta) is kept as w = c (1 + i t), c = cos(theta), t = -tan(theta): applying (1 + i t) costs two
// FMAs instead of the four operations of a complex multiply, and the real factor c is never applied: an element stored with a
// pending scale sigma (true value = sigma x stored value) is consumed by a DFT whose butterflies absorb the scales as RATIOS --
//     a + W b  with pending (sa, sb)   ->   (a + (sb/sa) W b)  with pending sa
// -- and every butterfly already ends in FMAs whose multiplier is a constant (1, 1/sqrt 2, cos pi/8), so the ratio replaces it for
// free (the products rho/sqrt 2 and rho cos(pi/8) cost four multiplies per 16-point DFT).  A DFT needs 8 + 4 + 2 + 1 = 15 ratios,
// which take the table space of the 15 real parts the twiddles no longer have.  Pending scales after a DFT equal the scale of the
// element in slot 0; the P2 / P2' DFTs multiply slot 0 by its scale first (n0, two multiplies), so that what they hand on carries
// only the next twiddle's c, and the P3 / P3' DFTs end with pending 1 because their slot-0 element carries the twiddle w^0 = 1:
// nothing is left to fold into the kernel spectrum or the post-factor.  theta is never reduced: c < 0 is fine, |t| <= 652 for
// N = 4096 (relative rounding is scale invariant); the one exact quarter turn per table (w = -i: c = 0) is marked by t = +inf and
// applied as the exact rotation (y, -x) with c = 1 by two predicated moves.
//
// per-thread table regions, 32 doubles each: [0,16) t_k (t_0 unused), [16,24) chunk A = {n0, r1[0..6]}, [24,32) chunk B =
// {r1[7], r2[0..3], r3[0..1], r4} = the ratios of the DFT that CONSUMES what was stored with the previous region's twiddles:
//   ST_TW1 : t of the P1 twiddles w_4096^{tau k1}            + ratios of this thread's P2  DFT (normalised)
//   ST_TW2 : t of the P2 twiddles w_256^{L l1}               + ratios of this thread's P3  DFT            (depends on L only)
//   ST_UT  : kernel spectrum, 16 complex
//   ST_TW1B: t of the P1' twiddles w_4096^{(H + 16 L) k1'}   + ratios of this thread's P2' DFT (normalised)
// and, uniform over a half-warp (shared memory): M[0][H][l1'] = t of the P2' twiddles w_256^{H l1'}, M[1][H][.] = ratios of P3'.
enum { ST_TW1 = 0, ST_TW2 = 1, ST_UT = 2, ST_TW1B = 3, ST_NTAB = 4 };

// Table provider interface (TB): issue<TABLE, SET>(chunk, buf) starts fetching chunk `chunk` (8 doubles = 4 complex) of the table
// into buffer `buf` (0/1), wait(buf) completes it, getd<SET>(TABLE, chunk, buf, i) / get<SET>(TABLE, chunk, buf, i) return double
// 8*chunk + i / complex 4*chunk + i.  The CUDA kernels read tensor memory (asynchronous tcgen05.ld, double buffered); the CPU
// emulation reads a plain array.

// (x + i y)(1 + i t); t = +inf marks the exact quarter turn w = -i
CPF_HD double2 st_mul_t(const double2 z, const double t) { return mk2(fma(-t, z.y, z.x), fma(t, z.x, z.y)); }
CPF_HD double2 st_mul_t8(const double2 z, const double t) {
  double2 r = st_mul_t(z, t);
  if (t > 1.7e308) r = mk2(z.y, -z.x);
  return r;
}

// butterfly (a, b) <- (a + m' W b, a - m' W b), W = exp(-2 pi i K/16); m is the ratio times the constant the plain butterfly has in
// its last FMAs (1 for K = 0, 4; 1/sqrt 2 for K = 2, 6; cos(pi/8) for odd K).  TOP: only a is needed.
template <int K, bool TOP = false>
CPF_HD void bflyr(double2& a, double2& b, const double m) {
  const double ax = a.x, ay = a.y, bx = b.x, by = b.y;
  double px, py;
  bool nx = false, ny = false;          // sign of the p-terms in the upper output
  if constexpr (K == 0) { px = bx; py = by; }
  else if constexpr (K == 4) { px = by; py = bx; ny = true; }
  else if constexpr (K == 2) { px = bx + by; py = by - bx; }
  else if constexpr (K == 6) { px = by - bx; py = bx + by; ny = true; }
  else if constexpr (K == 1) { px = fma(CPF_TAN_PI_8, by, bx); py = fma(-CPF_TAN_PI_8, bx, by); }
  else if constexpr (K == 7) { px = fma(-CPF_TAN_PI_8, by, bx); py = fma(CPF_TAN_PI_8, bx, by); nx = true; ny = true; }
  else if constexpr (K == 3) { px = fma(CPF_TAN_PI_8, bx, by); py = fma(CPF_TAN_PI_8, by, -bx); }
  else { px = fma(-CPF_TAN_PI_8, bx, by); py = fma(CPF_TAN_PI_8, by, bx); ny = true; }       // K == 5
  a.x = fma(nx ? -m : m, px, ax);
  a.y = fma(ny ? -m : m, py, ay);
  if constexpr (!TOP) {
    b.x = fma(nx ? m : -m, px, ax);
    b.y = fma(ny ? m : -m, py, ay);
  }
}

// In-register forward DFT of length 16 (decimation in time, slots as in dft_dit: w[bitrev(j)] = x[j]) of elements with pending
// scales, given the 16 table doubles A[8] = {n0, r1[0..6]}, B[8] = {r1[7], r2[0..3], r3[0..1], r4} through ga(i) / gb(i).
// NORM: slot 0 is multiplied by n0 first.  HALF_OUT: only w[k], k < 8, are valid on exit.
template <bool NORM, bool HALF_OUT, class GA, class GB, class HOOK>
CPF_HD void dft16_ratio(double2 (&w)[16], const GA& ga, const GB& gb, const HOOK& after_a) {
  if constexpr (NORM) { const double n0 = ga(0); w[0].x *= n0; w[0].y *= n0; }
#pragma unroll
  for (int i = 0; i < 7; ++i) bflyr<0>(w[2 * i], w[2 * i + 1], ga(i + 1));
  after_a();                       // chunk A is dead: its buffer can take the next table
  bflyr<0>(w[14], w[15], gb(0));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double r = gb(1 + j);
    bflyr<0>(w[4 * j], w[4 * j + 2], r);
    bflyr<4>(w[4 * j + 1], w[4 * j + 3], r);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double r = gb(5 + j), rs = r * CPF_SQRT1_2;
    bflyr<0>(w[8 * j], w[8 * j + 4], r);
    bflyr<2>(w[8 * j + 1], w[8 * j + 5], rs);
    bflyr<4>(w[8 * j + 2], w[8 * j + 6], r);
    bflyr<6>(w[8 * j + 3], w[8 * j + 7], rs);
  }
  {
    const double r = gb(7), rs = r * CPF_SQRT1_2, rc = r * CPF_COS_PI_8;
    bflyr<0, HALF_OUT>(w[0], w[8], r); bflyr<1, HALF_OUT>(w[1], w[9], rc); bflyr<2, HALF_OUT>(w[2], w[10], rs); bflyr<3, HALF_OUT>(w[3], w[11], rc);
    bflyr<4, HALF_OUT>(w[4], w[12], r); bflyr<5, HALF_OUT>(w[5], w[13], rc); bflyr<6, HALF_OUT>(w[6], w[14], rs); bflyr<7, HALF_OUT>(w[7], w[15], rc);
  }
}

// Ratio DFT with chunks 2 (A, buffer 0) and 3 (B, buffer 1) of TABLE, which the caller has issued; chunks 0 and 1 of NEXT (the t of
// the twiddles applied next, or the first two chunks of the kernel spectrum) are issued into the buffers as they fall free.
template <class TB, int TABLE, bool NORM, int NEXT, int SET>
CPF_HD void st_dft_ratio(TB& tb, double2 (&w)[16]) {
  tb.wait(0);
  tb.wait(1);
  dft16_ratio<NORM, false>(w, [&](const int i) { return tb.template getd<SET>(TABLE, 2, 0, i); }, [&](const int i) { return tb.template getd<SET>(TABLE, 3, 1, i); },
                           [&]() { if constexpr (NEXT >= 0) tb.template issue<(NEXT >= 0 ? NEXT : 0), SET>(0, 0); });
  if constexpr (NEXT >= 0) tb.template issue<(NEXT >= 0 ? NEXT : 0), SET>(1, 1);
}
// (1 + i t_k), k = 1..15, with chunks 0 (buffer 0) and 1 (buffer 1) of TABLE, already issued
template <class TB, int TABLE, int SET>
CPF_HD void st_apply_t(TB& tb, double2 (&w)[16]) {
  tb.wait(0);
#pragma unroll
  for (int k = 1; k < 8; ++k) w[k] = st_mul_t(w[k], tb.template getd<SET>(TABLE, 0, 0, k));
  tb.wait(1);
  w[8] = st_mul_t8(w[8], tb.template getd<SET>(TABLE, 1, 1, 0));
#pragma unroll
  for (int k = 9; k < 16; ++k) w[k] = st_mul_t(w[k], tb.template getd<SET>(TABLE, 1, 1, k - 8));
}
// kernel spectrum: plain complex multiplies, 4 chunks of 4 complex; chunks 0, 1 already issued into buffers 0, 1; chunks 0, 1 of NEXT
// are issued when the buffers fall free
template <class TB, int NEXT, int SET>
CPF_HD void st_apply_ut(TB& tb, double2 (&w)[16]) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    tb.wait(ch & 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[4 * ch + i] = cmul(w[4 * ch + i], tb.template get<SET>(ST_UT, ch, ch & 1, i));
    if (ch < 2) tb.template issue<ST_UT, SET>(ch + 2, ch & 1);
    else tb.template issue<NEXT, SET>(ch - 2, ch & 1);
  }
}

// ---- P1: v[r], r < 8 = element tau + 256 r of the (rotated) input window, pre-factor applied; r >= 8 are zero ----
template <class TB, int SET = 0>
CPF_HD void st_p1_compute(const double2 (&v)[8], double2 (&w)[16], TB& tb) {
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) w[bitrev(n1, 4)] = v[n1];
  tb.template issue<ST_TW1, SET>(0, 0);
  tb.template issue<ST_TW1, SET>(1, 1);
  dft_dit<16, true, false>(w);
  st_apply_t<TB, ST_TW1, SET>(tb, w);
}
CPF_HD void st_p1_store(const int tau, const double2 (&w)[16], double2* S) {
  double2* col = S + st_slot(0, tau >> 4, tau & 15);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) col[k1 * ST_RS] = w[k1];
}

// ---- P2 / P2': row H, stride 17 ----
CPF_HD void st_row_load(const int tau, const double2* S, double2 (&w)[16]) {
  const double2* row = S + st_slot(tau >> 4, 0, tau & 15);
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) w[bitrev(m1, 4)] = row[17 * m1];
}
CPF_HD void st_row_store(const int tau, const double2 (&w)[16], double2* S) {
  double2* row = S + st_slot(tau >> 4, 0, tau & 15);
#pragma unroll
  for (int l1 = 0; l1 < 16; ++l1) row[17 * l1] = w[l1];
}
// P2' twiddles and P3' ratios, uniform over a half-warp: Mt points at t of w_256^{H l1}, l1 = 0..15, Mr at the 16 ratio doubles of row H
CPF_HD void st_p2b_twiddle(double2 (&w)[16], const double* Mt) {
  const double2* t2 = reinterpret_cast<const double2*>(Mt);       // 16-byte loads of (t[2 i], t[2 i + 1])
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double2 tt = t2[i];
    if (i > 0) w[2 * i] = i == 4 ? st_mul_t8(w[8], tt.x) : st_mul_t(w[2 * i], tt.x);
    w[2 * i + 1] = st_mul_t(w[2 * i + 1], tt.y);
  }
}

// ---- P3, kernel multiply, P1': the thread's own 16 contiguous slots ----
CPF_HD void st_own_load(const int tau, const double2* S, double2 (&v)[16]) {
  const double2* own = S + st_slot(tau >> 4, tau & 15, 0);
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = own[m2];
}
CPF_HD void st_own_store(const int tau, const double2 (&w)[16], double2* S) {
  double2* own = S + st_slot(tau >> 4, tau & 15, 0);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) own[k1] = w[k1];
}
template <class TB, int SET = 0>
CPF_HD void st_p3_compute(double2 (&v)[16], double2 (&w)[16], TB& tb) {   // ratio chunks of ST_TW2 issued by the caller
  st_dft_ratio<TB, ST_TW2, false, ST_UT, SET>(tb, v);
  st_apply_ut<TB, ST_TW1B, SET>(tb, v);
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) w[bitrev(n1, 4)] = v[n1];
  dft_dit<16, false, false>(w);
  st_apply_t<TB, ST_TW1B, SET>(tb, w);
}

// ---- P3': column tau of every row; v[l2'], l2' < 8 = element tau + 256 l2' of the output window ----
CPF_HD void st_col_load(const int tau, const double2* S, double2 (&v)[16]) {
  const double2* col = S + st_slot(0, tau >> 4, tau & 15);
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = col[m2 * ST_RS];
}
CPF_HD void st_p3b_compute(double2 (&v)[16], const double* Mr) {
  const double2* r2 = reinterpret_cast<const double2*>(Mr);
  double2 q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = r2[i];
  dft16_ratio<false, true>(v, [&](const int i) { return (i & 1) ? q[i >> 1].y : q[i >> 1].x; },
                           [&](const int i) { return (i & 1) ? q[4 + (i >> 1)].y : q[4 + (i >> 1)].x; }, []() {});
}

// ---- whole phases (one column set per thread) ----
template <class TB>
CPF_HD void st_p1(const int tau, const double2 (&v)[8], double2* S, TB& tb) {
  double2 w[16];
  st_p1_compute(v, w, tb);
  st_p1_store(tau, w, S);
}
template <class TB>
CPF_HD void st_p2(const int tau, double2* S, TB& tb) {
  double2 w[16];
  tb.template issue<ST_TW1, 0>(2, 0);
  tb.template issue<ST_TW1, 0>(3, 1);
  st_row_load(tau, S, w);
  st_dft_ratio<TB, ST_TW1, true, ST_TW2, 0>(tb, w);
  st_apply_t<TB, ST_TW2, 0>(tb, w);
  st_row_store(tau, w, S);
}
template <class TB>
CPF_HD void st_p3_mul_p1(const int tau, double2* S, TB& tb) {
  double2 v[16], w[16];
  tb.template issue<ST_TW2, 0>(2, 0);
  tb.template issue<ST_TW2, 0>(3, 1);
  st_own_load(tau, S, v);
  st_p3_compute(v, w, tb);
  st_own_store(tau, w, S);
}
// M: shared [2][16][16] doubles (see above)
template <class TB>
CPF_HD void st_p2b(const int tau, double2* S, TB& tb, const double* M) {
  double2 w[16];
  tb.template issue<ST_TW1B, 0>(2, 0);
  tb.template issue<ST_TW1B, 0>(3, 1);
  st_row_load(tau, S, w);
  st_dft_ratio<TB, ST_TW1B, true, -1, 0>(tb, w);
  st_p2b_twiddle(w, M + 16 * (tau >> 4));
  st_row_store(tau, w, S);
}
CPF_HD void st_p3b(const int tau, double2 (&v)[16], const double2* S, const double* M) {
  st_col_load(tau, S, v);
  st_p3b_compute(v, M + 256 + 16 * (tau >> 4));
}

// ---- host side: the table values (used by cpf_fftlog.cu and by tests/emul/emul_stream.cpp) -----------------------------------
struct StScaled { double c, t; };
// w = exp(-2 pi i num/den) = c (1 + i t); the exact quarter turns are marked (c = 1, t = +-inf)
inline StScaled st_scaled_root(long long num, const long long den) {
  num %= den;
  StScaled r;
  if (4 * num == den) { r.c = 1.; r.t = HUGE_VAL; return r; }             // w = -i
  if (4 * num == 3 * den) { r.c = 1.; r.t = -HUGE_VAL; return r; }        // w = +i (does not occur for the index ranges of this kernel)
  if (num == 0) { r.c = 1.; r.t = 0.; return r; }
  if (2 * num == den) { r.c = -1.; r.t = 0.; return r; }
  const long double ang = -2.0L * acosl(-1.0L) * (long double)num / (long double)den;
  r.c = (double)cosl(ang);
  r.t = (double)tanl(ang);
  return r;
}
// the 16 ratio doubles {n0, r1[0..6]}, {r1[7], r2[0..3], r3[0..1], r4} of a DFT whose input j carries the pending scale sigma[j]
inline void st_ratio_table(const double (&sigma)[16], const bool norm, double* out) {
  long double s[16];
  for (int slot = 0; slot < 16; ++slot) s[slot] = sigma[bitrev(slot, 4)];
  out[0] = norm ? (double)s[0] : 1.;
  if (norm) s[0] = 1.L;
  for (int i = 0; i < 8; ++i) (i < 7 ? out[1 + i] : out[8]) = (double)(s[2 * i + 1] / s[2 * i]);
  for (int j = 0; j < 4; ++j) out[9 + j] = (double)(s[4 * j + 2] / s[4 * j]);
  for (int j = 0; j < 2; ++j) out[13 + j] = (double)(s[8 * j + 4] / s[8 * j]);
  out[15] = (double)(s[8] / s[0]);
}
// all batch-invariant tables of the stream kernel: tw [3][32][256] doubles (regions ST_TW1, ST_TW2, ST_TW1B in this order; double
// e of thread tau at tw[(region * 32 + e) * 256 + tau]) and M [2][16][16]
inline void st_build_tables(double* tw, double* M) {
  const int N = 4096, T = 256;
  for (int tau = 0; tau < T; ++tau) {
    const int H = tau >> 4, L = tau & 15;
    double sig[16], rat[16];
    auto put = [&](const int region, const int e, const double v) { tw[((size_t)region * 32 + e) * T + tau] = v; };
    for (int k = 0; k < 16; ++k) {
      put(0, k, st_scaled_root((long long)tau * k, N).t);                         // P1 : w_4096^{tau k1}
      put(1, k, st_scaled_root(L * k, 256).t);                                    // P2 : w_256^{L l1}
      put(2, k, st_scaled_root((long long)(H + 16 * L) * k, N).t);                // P1': w_4096^{(H + 16 L) k1'}
    }
    // P2 of thread (k1, m2) = (H, L): input m1 comes from thread 16 m1 + L, P1 output k1 = H
    for (int m1 = 0; m1 < 16; ++m1) sig[m1] = st_scaled_root((long long)(16 * m1 + L) * H, N).c;
    st_ratio_table(sig, true, rat);
    for (int e = 0; e < 16; ++e) put(0, 16 + e, rat[e]);
    // P3 of thread (k1, l1) = (H, L): input m2 comes from thread (H, m2), P2 output l1 = L (P2 is normalised)
    for (int m2 = 0; m2 < 16; ++m2) sig[m2] = st_scaled_root(m2 * L, 256).c;
    st_ratio_table(sig, false, rat);
    for (int e = 0; e < 16; ++e) put(1, 16 + e, rat[e]);
    // P2' of thread (m2', k1') = (H, L): input m1' comes from thread (H, m1'), P1' output k1' = L
    for (int m1 = 0; m1 < 16; ++m1) sig[m1] = st_scaled_root((long long)(H + 16 * m1) * L, N).c;
    st_ratio_table(sig, true, rat);
    for (int e = 0; e < 16; ++e) put(2, 16 + e, rat[e]);
  }
  for (int H = 0; H < 16; ++H) {
    double sig[16];
    for (int l = 0; l < 16; ++l) M[16 * H + l] = st_scaled_root(H * l, 256).t;      // P2': w_256^{H l1'}
    // P3' of thread (l1', k1') = (H, L): input m2' comes from thread (m2', L), P2' output l1' = H (P2' is normalised)
    for (int m2 = 0; m2 < 16; ++m2) sig[m2] = st_scaled_root(m2 * H, 256).c;
    st_ratio_table(sig, false, M + 256 + 16 * H);
  }
}
}  // namespace cpf
