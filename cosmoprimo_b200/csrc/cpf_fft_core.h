// cpf_fft_core.h — register-level FFT building blocks shared by the sm_100a kernels and by the CPU thread-emulation
// test (tests/emul).  Everything here is __host__ __device__ straight-line fp64 code: no memory traffic except
// through the pointers the caller passes, no synchronisation.
//
// Design: a length-N complex FFT (N = 256*R1, R1 in {4,8,16}) is done by T = 16*R1 threads, each
// holding 16 complex values in registers, in three register passes (radix R1, 16, 16) separated by two shared-memory
// exchanges.  Thread t holds elements t + T*r (r = 0..15) on entry AND on exit (natural order), so two FFTs chain
// without an exchange in between.  All DFTs are forward (e^{-2 pi i jk/N}); FFTLog's `irfft(conj(.))`
// (cosmoprimo/fftlog.py:544) is a second forward transform — see cpf_fftlog.cu.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CPF_HD __host__ __device__ __forceinline__
#define CPF_HDC __host__ __device__ constexpr
#else
#define CPF_HD inline
#define CPF_HDC constexpr
#ifndef CPF_HAVE_DOUBLE2
#define CPF_HAVE_DOUBLE2
struct double2 { double x, y; };
#endif
#endif

// read-only (non-coherent) load for the batch-invariant tables; plain load in the CPU emulation
#if defined(__CUDA_ARCH__)
#define CPF_LDG(ptr) __ldg(ptr)
#else
#define CPF_LDG(ptr) (*(ptr))
#endif

namespace cpf {

CPF_HD double2 mk2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

CPF_HD double2 cmul(const double2 a, const double2 b) {
  return mk2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

CPF_HD double2 cmul_conj(const double2 a, const double2 b) {   // a * conj(b)
  return mk2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}

// constants of the 16-point DFT
#define CPF_SQRT1_2 0.70710678118654752440
#define CPF_COS_PI_8 0.92387953251128675613
#define CPF_TAN_PI_8 0.41421356237309504880

// Radix-2 DIT butterfly with the constant twiddle w = exp(-2 pi i K/16):  (a, b) <- (a + w b, a - w b).
// General twiddles use the factored form w = C (1 + i t) so that every output is one FMA (6 FMAs per butterfly
// instead of 4 + 4); |t| <= tan(pi/8), so no precision is lost.
template <int K>
CPF_HD void bfly(double2& a, double2& b) {
  static_assert(K >= 0 && K < 8, "twiddle index");
  const double ax = a.x, ay = a.y, bx = b.x, by = b.y;
  if constexpr (K == 0) {
    a.x = ax + bx; a.y = ay + by; b.x = ax - bx; b.y = ay - by;
  } else if constexpr (K == 4) {           // w = -i : w b = (by, -bx)
    a.x = ax + by; a.y = ay - bx; b.x = ax - by; b.y = ay + bx;
  } else if constexpr (K == 2) {           // w = (1 - i)/sqrt2 : w b = c (bx + by, by - bx)
    const double px = bx + by, py = by - bx;
    a.x = fma(CPF_SQRT1_2, px, ax); a.y = fma(CPF_SQRT1_2, py, ay);
    b.x = fma(-CPF_SQRT1_2, px, ax); b.y = fma(-CPF_SQRT1_2, py, ay);
  } else if constexpr (K == 6) {           // w = (-1 - i)/sqrt2 : w b = c (by - bx, -(bx + by))
    const double px = by - bx, py = bx + by;
    a.x = fma(CPF_SQRT1_2, px, ax); a.y = fma(-CPF_SQRT1_2, py, ay);
    b.x = fma(-CPF_SQRT1_2, px, ax); b.y = fma(CPF_SQRT1_2, py, ay);
  } else if constexpr (K == 1) {           // w = C (1, -T) : w b = C (bx + T by, by - T bx)
    const double px = fma(CPF_TAN_PI_8, by, bx), py = fma(-CPF_TAN_PI_8, bx, by);
    a.x = fma(CPF_COS_PI_8, px, ax); a.y = fma(CPF_COS_PI_8, py, ay);
    b.x = fma(-CPF_COS_PI_8, px, ax); b.y = fma(-CPF_COS_PI_8, py, ay);
  } else if constexpr (K == 7) {           // w = -C (1, T) : w b = -C (bx - T by, by + T bx)
    const double px = fma(-CPF_TAN_PI_8, by, bx), py = fma(CPF_TAN_PI_8, bx, by);
    a.x = fma(-CPF_COS_PI_8, px, ax); a.y = fma(-CPF_COS_PI_8, py, ay);
    b.x = fma(CPF_COS_PI_8, px, ax); b.y = fma(CPF_COS_PI_8, py, ay);
  } else if constexpr (K == 3) {           // w = C (T, -1) : w b = C (T bx + by, T by - bx)
    const double px = fma(CPF_TAN_PI_8, bx, by), py = fma(CPF_TAN_PI_8, by, -bx);
    a.x = fma(CPF_COS_PI_8, px, ax); a.y = fma(CPF_COS_PI_8, py, ay);
    b.x = fma(-CPF_COS_PI_8, px, ax); b.y = fma(-CPF_COS_PI_8, py, ay);
  } else {                       // K == 5 : w = C (-T, -1) : w b = C (by - T bx, -(bx + T by))
    const double px = fma(-CPF_TAN_PI_8, bx, by), py = fma(CPF_TAN_PI_8, by, bx);
    a.x = fma(CPF_COS_PI_8, px, ax); a.y = fma(-CPF_COS_PI_8, py, ay);
    b.x = fma(-CPF_COS_PI_8, px, ax); b.y = fma(CPF_COS_PI_8, py, ay);
  }
}

// Same butterfly when only the upper output a + w b is wanted (output pruning of the last stage).
template <int K>
CPF_HD void bfly_top(double2& a, const double2 b) {
  double2 bb = b;
  // the compiler drops the dead half of bfly<K>; spelled out here only for K where that saves the p-terms too
  bfly<K>(a, bb);
}

CPF_HDC int bitrev(int v, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
  return r;
}

// In-register forward DFT of length R (4, 8 or 16), decimation in time.
//   on entry  w[bitrev(j)] = x[j]   (the caller loads straight into the bit-reversed slot: indices are compile-time)
//   on exit   w[k] = sum_j x[j] exp(-2 pi i jk/R)
// HALF_IN : x[j] = 0 for j >= R/2 (odd slots are ignored and overwritten) — FFTLog's zero padding.
// HALF_OUT: only w[k], k < R/2, are valid on exit — FFTLog's output crop.
template <int R, bool HALF_IN, bool HALF_OUT, int M>
CPF_HD void dft_dit(double2 (&w)[M]) {   // only w[0..R) are touched
  static_assert(R == 4 || R == 8 || R == 16, "radix");
  static_assert(M >= R, "array too short");
  // stage h = 1
#pragma unroll
  for (int j = 0; j < R; j += 2) {
    if constexpr (HALF_IN) w[j + 1] = w[j];
    else bfly<0>(w[j], w[j + 1]);
  }
  // stage h = 2 : twiddles w4^i = w16^{4 i}
  if constexpr (R == 4 && HALF_OUT) {
    bfly_top<0>(w[0], w[2]); bfly_top<4>(w[1], w[3]);
  }
  else {
#pragma unroll
    for (int b = 0; b < R; b += 4) {
      bfly<0>(w[b], w[b + 2]);
      bfly<4>(w[b + 1], w[b + 3]);
    }
  }
  if constexpr (R == 4) return;
  // stage h = 4 : twiddles w8^i = w16^{2 i}
  if constexpr (R == 8 && HALF_OUT) {
    bfly_top<0>(w[0], w[4]); bfly_top<2>(w[1], w[5]); bfly_top<4>(w[2], w[6]); bfly_top<6>(w[3], w[7]);
  } else if constexpr (R >= 8) {
#pragma unroll
    for (int b = 0; b < R; b += 8) {
      bfly<0>(w[b], w[b + 4]);
      bfly<2>(w[b + 1], w[b + 5]);
      bfly<4>(w[b + 2], w[b + 6]);
      bfly<6>(w[b + 3], w[b + 7]);
    }
  }
  if constexpr (R == 8) return;
  // stage h = 8 : twiddles w16^i
  if constexpr (R == 16 && HALF_OUT) {
    bfly_top<0>(w[0], w[8]); bfly_top<1>(w[1], w[9]); bfly_top<2>(w[2], w[10]); bfly_top<3>(w[3], w[11]);
    bfly_top<4>(w[4], w[12]); bfly_top<5>(w[5], w[13]); bfly_top<6>(w[6], w[14]); bfly_top<7>(w[7], w[15]);
  } else if constexpr (R == 16) {
    bfly<0>(w[0], w[8]); bfly<1>(w[1], w[9]); bfly<2>(w[2], w[10]); bfly<3>(w[3], w[11]);
    bfly<4>(w[4], w[12]); bfly<5>(w[5], w[13]); bfly<6>(w[6], w[14]); bfly<7>(w[7], w[15]);
  }
}

CPF_HDC int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// Geometry of the three-pass FFT.
template <int R1>
struct Geo {
  static constexpr int T = 16 * R1;        // threads per transform
  static constexpr int N = 256 * R1;       // transform length
  static constexpr int C = 16 / R1;        // radix-R1 columns per thread in pass 1
  static constexpr int RS = 257;           // shared-memory row stride in complex elements (odd => conflict-free)
  static constexpr int SMEM_ELEMS = R1 * RS;
  static constexpr int B1 = ilog2c(R1);
};

// ---- factored twiddles -------------------------------------------------------------------------------------------
// A pass needs w^k, k = 1..R-1, for one per-thread base w.  Loading all of them costs more L1 bandwidth (the
// kernel's scarcest resource) than fp64 issue slots, so only w^1, w^2, w^3, w^4, w^8, w^12 are loaded
// (table slots 0..5, each correctly rounded) and the rest are formed as w^(4c) * w^d: one extra rounding, 9 complex
// multiplies instead of 9 more 16-byte loads per thread and pass.
template <int R>
struct TwSet {
  double2 b[6];
  // base points at slot 0 for this thread, slots are `stride` elements apart
  CPF_HD void load(const double2* base, const int stride) {
#pragma unroll
    for (int s = 0; s < 6; ++s)
      if (s < 3 || (s == 3 && R > 4) || (s > 3 && R > 8)) b[s] = base[s * stride];
  }
  template <int K>
  CPF_HD double2 get() const {
    static_assert(K >= 1 && K < R, "twiddle power");
    constexpr int c = K >> 2, d = K & 3;
    if constexpr (c == 0) return b[d - 1];
    else if constexpr (d == 0) return b[2 + c];
    else return cmul(b[2 + c], b[d - 1]);
  }
};

template <int R, int K>
struct TwApply {   // S[k * stride] = w[k] * tw^k for k = K..R-1 (compile-time recursion keeps K a constant)
  template <typename TW>
  static CPF_HD void run(double2* dst, const int stride, const double2 (&w)[R], const TW& tw) {
    dst[K * stride] = cmul(w[K], tw.template get<K>());
    if constexpr (K + 1 < R) TwApply<R, K + 1>::run(dst, stride, w, tw);
  }
};

// ---- pass 1: radix-R1 over n1 (stride 256), twiddle w_N^{n2 k1}, scatter to S[k1][n2] ------------------------
// v[r] holds element t + T*r = 256*n1 + n2 with r = n1*C + c, n2 = t + T*c.
// tw1[s*256 + n2] = exp(-2 pi i n2 e_s / N), e_s = {1,2,3,4,8,12}.
template <int R1, bool HALF_IN>
CPF_HD void fft_pass1(const int t, const double2 (&v)[16], double2* S, const double2* tw1) {
  typedef Geo<R1> G;
#pragma unroll
  for (int c = 0; c < G::C; ++c) {
    const int n2 = t + G::T * c;
    TwSet<R1> tw;
    tw.load(tw1 + n2, 256);
    double2 w[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) w[bitrev(n1, G::B1)] = v[n1 * G::C + c];
    dft_dit<R1, HALF_IN, false>(w);
    S[n2] = w[0];
    TwApply<R1, 1>::run(S + n2, G::RS, w, tw);
  }
}

// ---- pass 2: thread (k1, m2) = (t/16, t%16): radix-16 over m1, twiddle w_256^{m2 l1}, write back in place ------
// tw2[s*16 + m2] = exp(-2 pi i m2 e_s / 256).
template <int R1>
CPF_HD void fft_pass2(const int t, double2* S, const double2* tw2) {
  typedef Geo<R1> G;
  const int k1 = t >> 4, m2 = t & 15;
  double2* row = S + k1 * G::RS + m2;
  TwSet<16> tw;
  tw.load(tw2 + m2, 16);
  double2 w[16];
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) w[bitrev(m1, 4)] = row[16 * m1];
  dft_dit<16, false, false>(w);
  row[0] = w[0];
  TwApply<16, 1>::run(row, 16, w, tw);
}

// ---- pass 3: thread (k1, l1) = (t%R1, t/R1): radix-16 over m2; v[l2] = X[t + T*l2] ----------------------------
template <int R1, bool HALF_OUT>
CPF_HD void fft_pass3(const int t, double2 (&v)[16], const double2* S) {
  typedef Geo<R1> G;
  const int k1 = t & (R1 - 1), l1 = t >> G::B1;
  const double2* row = S + k1 * G::RS + 16 * l1;
#pragma unroll
  for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = row[m2];
  dft_dit<16, false, HALF_OUT>(v);
}

}  // namespace cpf
