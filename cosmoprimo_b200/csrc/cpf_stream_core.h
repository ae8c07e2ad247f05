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

// per-thread tables, 16 complex entries each (entry 0 of the twiddle tables is 1 and is never read)
enum { ST_TW1 = 0, ST_TW2 = 1, ST_UT = 2, ST_TW1B = 3, ST_NTAB = 4 };

// Table provider interface (TB): issue<TABLE, SET>(chunk, buf) starts fetching entries 4*chunk..4*chunk+3 of the
// tables of column set SET into buffer `buf` (0/1), wait(buf) completes it, get<SET>(TABLE, chunk, buf, i) returns entry
// 4*chunk + i.  The CUDA kernels read tensor memory (asynchronous tcgen05.ld, double buffered); the CPU emulation
// reads a plain array.
template <class TB, int TABLE, bool SKIP0, int SET = 0>
CPF_HD void st_apply(TB& tb, double2 (&w)[16]) {
  tb.template issue<TABLE, SET>(0, 0);
  tb.wait(0);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (ch < 3) tb.template issue<TABLE, SET>(ch + 1, (ch + 1) & 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 4 * ch + i;
      if (!(SKIP0 && k == 0)) w[k] = cmul(w[k], tb.template get<SET>(TABLE, ch, ch & 1, i));
    }
    if (ch < 3) tb.wait((ch + 1) & 1);
  }
}

// ---- P1: v[r], r < 8 = element tau + 256 r of the (rotated) input window, pre-factor applied; r >= 8 are zero ----
template <class TB, int SET = 0>
CPF_HD void st_p1_compute(const double2 (&v)[8], double2 (&w)[16], TB& tb) {
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) w[bitrev(n1, 4)] = v[n1];
  dft_dit<16, true, false>(w);
  st_apply<TB, ST_TW1, true, SET>(tb, w);
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
template <class TB, int SET = 0>
CPF_HD void st_p2_compute(double2 (&w)[16], TB& tb) {
  dft_dit<16, false, false>(w);
  st_apply<TB, ST_TW2, true, SET>(tb, w);
}
// M16 points at w_256^{H l1}, l1 = 0..15 (the same for the 16 threads of a half-warp)
// (GLOBAL: the table is in global memory and is read through the read-only path)
template <bool GLOBAL = false>
CPF_HD void st_p2b_compute(double2 (&w)[16], const double2* M16) {
  dft_dit<16, false, false>(w);
#pragma unroll
  for (int l1 = 1; l1 < 16; ++l1) w[l1] = cmul(w[l1], GLOBAL ? CPF_LDG(M16 + l1) : M16[l1]);
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
CPF_HD void st_p3_compute(double2 (&v)[16], double2 (&w)[16], TB& tb) {
  dft_dit<16, false, false>(v);
  st_apply<TB, ST_UT, false, SET>(tb, v);
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) w[bitrev(n1, 4)] = v[n1];
  dft_dit<16, false, false>(w);
  st_apply<TB, ST_TW1B, true, SET>(tb, w);
}

// ---- P3': column tau of every row; v[l2'], l2' < 8 = element tau + 256 l2' of the output window ----
CPF_HD void st_col_load(const int tau, const double2* S, double2 (&v)[16]) {
  const double2* col = S + st_slot(0, tau >> 4, tau & 15);
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = col[m2 * ST_RS];
}
CPF_HD void st_p3b_compute(double2 (&v)[16]) { dft_dit<16, false, true>(v); }

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
  st_row_load(tau, S, w);
  st_p2_compute(w, tb);
  st_row_store(tau, w, S);
}
template <class TB>
CPF_HD void st_p3_mul_p1(const int tau, double2* S, TB& tb) {
  double2 v[16], w[16];
  st_own_load(tau, S, v);
  st_p3_compute(v, w, tb);
  st_own_store(tau, w, S);
}
CPF_HD void st_p2b(const int tau, double2* S, const double2* M) {
  double2 w[16];
  st_row_load(tau, S, w);
  st_p2b_compute(w, M + 16 * (tau >> 4));
  st_row_store(tau, w, S);
}
CPF_HD void st_p3b(const int tau, double2 (&v)[16], const double2* S) {
  st_col_load(tau, S, v);
  st_p3b_compute(v);
}

}  // namespace cpf
