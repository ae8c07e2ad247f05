// cpf_wallish.cu — Wallish2018 no-wiggle filter (cosmoprimo/bao_filter.py:361-431) and the orthonormal DST-II/III it is
// built on (scipy.fftpack.dst / idst, bao_filter.py:372, 412), on sm_100a.
//
// cpf_wallish2018 is ONE kernel, wallish_fused_kernel: one CTA of 256 threads per PAIR of spectra (columns of the reference layout,
// wavenumber along axis 0), two CTAs per SM, everything between the load of pklin and the store of pknow in ONE 68 KB shared-memory
// buffer and in registers (cpf_wallish_core.h):
//   log(k P) -> DST-II (one packed complex FFT-4096, the FFTLog register FFT) -> clamped-spline second derivatives of the even / odd
//   coefficients -> argmax boxes -> cut + re-spline -> DST-III -> exp(.)/k on 1e-2 < k < 1.5 -> spliced with the unfiltered spectrum
//   at k < 5e-4 and k > 2 -> clamped spline on those knots, evaluated at self.k -> blend with the Gaussian top-hat (:415-423).
// HBM traffic per spectrum: 8 * 4096 B of pklin in, 8 * nk B of pk in and of pknow out; no intermediate leaves the SM.
#include <math.h>
#include <mutex>
#include <string.h>
#include <vector>

#include "cpf_async.h"
#include "cpf_common.h"
#include "cpf_fastmath.h"
#include "cpf_fft_core.h"
#include "cpf_spline_core.h"
#include "cpf_wallish_core.h"
#include "cpf_wallish_final.h"

#ifndef CPF_WALLISH_L2_HINT
#define CPF_WALLISH_L2_HINT 256
#endif

namespace cpf {

struct WallishTables {
  int device;
  double2* tw1;    // [6,256] factored FFT twiddles, N = 4096
  double2* tw2;    // [6,16]
  double2* twd;    // [4096] exp(-i pi k / 2N)
  double* wtab;    // [32] Thomas pivots of the clamped uniform system
};

struct WallishArgs {
  const double* klin;     // [4096]
  const double* pklin;    // [4096, ld], or with lin_rows [ncols, 4096]: one contiguous row per spectrum (the layout cpf_spline_eval_t writes)
  int lin_rows;
  const double* pkout;    // [nk, ld]
  double* pknow;          // [nk, ld]
  long long ncols;
  long long ld;           // doubles between consecutive rows of pklin / pkout / pknow
  int vec;                // pairs of columns are 16-byte aligned in all three arrays
  int* boxes;             // [ncols, 4] or null
  unsigned long long* dbg; // lab: per-phase cycle totals (thread 0 of every CTA), or null
  const double2 *tw1, *tw2, *twd;
  const double* wtab;
  // final stage (cpf_wallish_final.h)
  int i0, i1, nl, nr, lz, rz, nmid, nc, nk, nrounds, slbase;
  WallishFinFac fc;       // elimination factors of the spliced spline
  const double* rklin;    // [4096] 1 / klin
  const int* slotT;
  const int* qstart;
  const int* qinfo;
  const double* qh;
  const double* qth;      // [nk] Gaussian top-hat at self.k (:425-431)
};

// shared-memory carve-up
struct WallishSmem {
  double2* B;     // FFT exchange buffer / spectrum in natural order / DST coefficients (de-interleaved + padded) / knots + slope slots of the final spline
  double2* E;     // [256] chunk results of the two-step eliminations (value with zero inflow), forward
  double2* Eb;    // ... backward: the SAME memory (WallishCtaSync separates the last read of E from the write of Eb)
  double* Mf;     // [256] ... and the factors the inflow is multiplied with
  double* Mb;     // the same memory as Mf
  double* red;    // [32]
  int* redi;      // [32]
  int* box;       // [8]
  WallishGap* gaps;   // [4]
  double* wtab;   // [32] Thomas pivots (copied from global memory once: they sit in the dependency chain of the first chunks)
  __device__ explicit WallishSmem(double2* base) {
    B = base; E = base + WallishGeo::BUF; Eb = E;
    Mf = reinterpret_cast<double*>(E + 256);
    Mb = Mf;
    red = Mf + 256;
    wtab = red + 32;
    gaps = reinterpret_cast<WallishGap*>(wtab + 32);
    redi = reinterpret_cast<int*>(gaps + 4);
    box = redi + 32;
  }
};
struct WallishCtaSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
static constexpr size_t kWallishSmemBytes = ((size_t)WallishGeo::BUF + 256) * sizeof(double2) + (256 + 32 + 32) * sizeof(double) + 4 * sizeof(WallishGap) + 40 * sizeof(int);
static_assert(3 * (kWallishSmemBytes + 1024) <= 228 * 1024, "three CTAs per SM");

__device__ __forceinline__ void fft4096(const int t, double2 (&v)[16], double2* S, const double2* tw1, const double2* tw2) {
  fft_pass1<16, false>(t, v, S, tw1);
  __syncthreads();
  fft_pass2<16>(t, S, tw2);
  __syncthreads();
  fft_pass3<16, false>(t, v, S);
}

// DST-II (orthonormal) of the two packed real sequences whose Makhoul-permuted samples are in v; coefficients in B (padded layout)
__device__ __forceinline__ void dst2_in_smem(const int t, double2 (&v)[16], double2* B, const double2* tw1, const double2* tw2, const double2* twd) {
  fft4096(t, v, B, tw1, tw2);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) B[t + 256 * r] = v[r];
  __syncthreads();
  wallish_dst2_coef(t, v, B, twd);
  __syncthreads();
  wallish_dst2_store(t, v, B);
  __syncthreads();
}

// the part of dst2_in_smem after the FFT
__device__ __forceinline__ void dst2_post_in_smem(const int t, double2 (&v)[16], double2* B, const double2* twd) {
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) B[t + 256 * r] = v[r];
  __syncthreads();
  wallish_dst2_coef(t, v, B, twd);
  __syncthreads();
  wallish_dst2_store(t, v, B);
  __syncthreads();
}

// Makhoul input position: FFT sample n takes x'[j], x' = (-1)^j x
__device__ __forceinline__ int makhoul_src(const int n, double& sign) {
  if (n < WallishGeo::N / 2) { sign = 1.; return 2 * n; }
  sign = -1.;
  return 2 * (WallishGeo::N - 1 - n) + 1;
}

// the two columns of a pair at row `row` of a [rows, ld] array
__device__ __forceinline__ double2 load_pair(const double* base, const long long ld, const long long row, const long long col0, const bool has1, const int vec) {
  const double* p = base + row * ld + col0;
  if (vec && has1) {
    // a pair is 16 bytes of a row; the CTAs that run at the same time work on neighbouring pairs, so the whole 256-byte stretch of the row
    // is asked for at once (L2 prefetch size): 8 x fewer, 8 x longer DRAM accesses than one 32-byte sector per pair
    double2 r;
#if CPF_WALLISH_L2_HINT == 256
    asm volatile("ld.global.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
#elif CPF_WALLISH_L2_HINT == 128
    asm volatile("ld.global.L2::128B.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
#else
    r = __ldcs(reinterpret_cast<const double2*>(p));
#endif
    return r;
  }
  const double a = __ldcs(p);
  return mk2(a, has1 ? __ldcs(p + 1) : a);
}

// warp-level argmax of the chunk candidates of both columns; lane 0 of warp w (sequence parity h = w / 4) leaves the
// warp's best of column col in red / redi [2 w + col]
__device__ __forceinline__ void wallish_best_reduce(const int t, WallishBest b, double* red, int* redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ovx = __shfl_xor_sync(0xffffffffu, b.vx, o), ovy = __shfl_xor_sync(0xffffffffu, b.vy, o);
    const int oix = __shfl_xor_sync(0xffffffffu, b.ix, o), oiy = __shfl_xor_sync(0xffffffffu, b.iy, o);
    wallish_best_merge(b.vx, b.ix, ovx, oix);
    wallish_best_merge(b.vy, b.iy, ovy, oiy);
  }
  if ((t & 31) == 0) {
    const int w = t >> 5;
    red[2 * w] = b.vx; redi[2 * w] = b.ix;
    red[2 * w + 1] = b.vy; redi[2 * w + 1] = b.iy;
  }
}

// sequence q = 2 h + col: merge the four warps of parity h
__device__ __forceinline__ int wallish_best_final(const int q, const double* red, const int* redi) {
  const int h = q >> 1, col = q & 1;
  double v = 0.;
  int i = -1;
#pragma unroll
  for (int w = 4 * h; w < 4 * h + 4; ++w) wallish_best_merge(v, i, red[2 * w + col], redi[2 * w + col]);
  return i;
}

// ---- phases of the fused kernel ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_pair(double* base, const long long ld, const long long row, const long long col0, const bool has1, const int vec,
                                           const double2 val) {
  double* dst = base + row * ld + col0;
  if (vec && has1) __stcs(reinterpret_cast<double2*>(dst), val);
  else { dst[0] = val.x; if (has1) dst[1] = val.y; }
}

// output wavenumbers that are knots of the spliced spline (self.k < 5e-4, self.k > 2): it returns pk there, wiggles = 1, pknow = pk
__device__ __forceinline__ void wallish_copy_edges(const WallishArgs& a, const int t, const long long col0, const bool has1) {
  for (int c = t; c < a.nl + a.nr; c += WallishGeo::T) {
    const int q = c < a.nl ? c : a.nk - a.nr + (c - a.nl);
    store_pair(a.pknow, a.ld, q, col0, has1, a.vec, load_pair(a.pkout, a.ld, q, col0, has1, a.vec));
  }
}

// v[r] = sign * log(k P) in Makhoul order, straight from the reference layout (16 bytes per row and pair)      (bao_filter.py:371)
// ROLL: the samples land in the (free) buffer by asynchronous 16-byte copies, the logarithms are taken in place by a rolled loop
template <bool ROLL>
__device__ __forceinline__ void wallish_load_log(const WallishArgs& a, const WallishSmem& sm, const int t, const long long col0, const bool has1,
                                                 double2 (&v)[16], uint64_t* bar, unsigned& bar_parity) {
  typedef WallishGeo G;
  if (a.lin_rows) {
    // one contiguous 32 KB row per spectrum: two bulk copies per pair instead of 4096 16-byte requests through the L1 tag stage; row a lands in
    // raw[0, 4096), row b in raw[4096, 8192); every thread takes the logarithms of its own 16 + 16 samples in place, then picks them up
    double* raw = reinterpret_cast<double*>(sm.B);
    if (t == 0) {
      mbar_expect_tx(bar, has1 ? 2u * G::N * 8u : G::N * 8u);
      bulk_g2s(raw, a.pklin + col0 * G::N, G::N * 8u, bar);
      if (has1) bulk_g2s(raw + G::N, a.pklin + (col0 + 1) * G::N, G::N * 8u, bar);
    }
    wallish_copy_edges(a, t, col0, has1);
    mbar_wait(bar, bar_parity);
    bar_parity ^= 1u;
#pragma unroll 1
    for (int r4 = 0; r4 < 16; r4 += 4) {                // four samples per step: their k and shared-memory loads are in flight together
      double k[4], xa[4], xb[4];
      int j[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double sign;
        j[i] = makhoul_src(t + 256 * (r4 + i), sign);
        k[i] = __ldg(a.klin + j[i]);
        xa[i] = raw[j[i]];
        xb[i] = has1 ? raw[G::N + j[i]] : xa[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double sign = t + 256 * (r4 + i) < G::N / 2 ? 1. : -1.;
        raw[j[i]] = sign * fast_log(k[i] * xa[i]);
        raw[G::N + j[i]] = sign * fast_log(k[i] * xb[i]);
      }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = makhoul_src(t + 256 * r, sign);
      v[r] = mk2(raw[j], raw[G::N + j]);
    }
    __syncthreads();                                     // pass 1 of the FFT scatters into other threads' slots
    return;
  }
  if (!ROLL) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = makhoul_src(t + 256 * r, sign);
      v[r] = load_pair(a.pklin, a.ld, j, col0, has1, a.vec);
    }
    wallish_copy_edges(a, t, col0, has1);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = makhoul_src(t + 256 * r, sign);
      const double k = __ldg(a.klin + j);
      v[r] = mk2(sign * fast_log(k * v[r].x), sign * fast_log(k * v[r].y));
    }
    return;
  }
  if (a.vec && has1) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int n = t + 256 * r;
      double sign;
      const int j = makhoul_src(n, sign);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sm.B + n);
#if CPF_WALLISH_L2_HINT == 256
      asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst), "l"(a.pklin + (long long)j * a.ld + col0) : "memory");
#elif CPF_WALLISH_L2_HINT == 128
      asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst), "l"(a.pklin + (long long)j * a.ld + col0) : "memory");
#else
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(a.pklin + (long long)j * a.ld + col0) : "memory");
#endif
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
#pragma unroll 1
    for (int r = 0; r < 16; ++r) {
      const int n = t + 256 * r;
      double sign;
      const int j = makhoul_src(n, sign);
      sm.B[n] = load_pair(a.pklin, a.ld, j, col0, has1, 0);
    }
  }
  wallish_copy_edges(a, t, col0, has1);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
  for (int r4 = 0; r4 < 16; r4 += 4) {                  // four samples per step: their k and shared-memory loads are in flight together
    double k[4];
    double2 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = t + 256 * (r4 + i);
      double sign;
      k[i] = __ldg(a.klin + makhoul_src(n, sign));
      x[i] = sm.B[n];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = t + 256 * (r4 + i);
      const double sign = n < G::N / 2 ? 1. : -1.;
      sm.B[n] = mk2(sign * fast_log(k[i] * x[i].x), sign * fast_log(k[i] * x[i].y));
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = sm.B[t + 256 * r];
  __syncthreads();                                     // pass 1 of the FFT scatters into other threads' slots
}

// between the two FFTs: DST-II coefficients -> second derivatives -> boxes -> cut + re-spline -> DST-III input in v
__device__ __forceinline__ void wallish_middle(const WallishArgs& a, const WallishSmem& sm, const int t, const long long col0, const bool has1,
                                               double2 (&v)[16], double2 (&d)[16]) {
  typedef WallishGeo G;
  dst2_post_in_smem(t, v, sm.B, a.twd);                                                      // :372
  // second derivatives of the clamped splines through the even / odd coefficients           (:377-382); chunks live in registers
  wallish_forward_local(t, sm.B, d, sm.E, sm.Mf, sm.wtab);
  __syncthreads();
  wallish_forward_fix_backward_local(t, d, sm.E, sm.Mf, sm.Eb, sm.Mb, sm.wtab, WallishCtaSync());
  __syncthreads();
  const WallishBest chunk = wallish_backward_dd(t, sm.B, d, sm.Eb, sm.Mb, sm.wtab);
  // boxes (:392-395): argmax over [20, H-20), then over [first + 5, H-20); per-chunk maxima come out of the backward pass,
  // warps reduce them with shuffles, thread q < 4 merges the four warps of its sequence
  wallish_best_reduce(t, chunk, sm.red, sm.redi);
  __syncthreads();
  if (t < 4) sm.box[2 * t] = wallish_best_final(t, sm.red, sm.redi);
  __syncthreads();
  {
    const int h = t >> 7;
    const WallishBest cand = wallish_chunk_candidate(t, d, sm.box[4 * h] + G::MARGIN_SECOND, sm.box[4 * h + 2] + G::MARGIN_SECOND, chunk);
    wallish_best_reduce(t, cand, sm.red, sm.redi);     // red / redi were consumed before the previous barrier
  }
  __syncthreads();
  if (t < 4) {
    const int col = t & 1, h = t >> 1;
    const int amax = sm.box[2 * t], bmax = wallish_best_final(t, sm.red, sm.redi);
    const int b0 = amax + G::OFF_LO, b1 = (bmax < 0 ? G::H : bmax + G::OFF_HI);   // empty second range: treated as "to the end"
    sm.box[2 * t] = b0;
    sm.box[2 * t + 1] = b1;
    if (a.boxes && (col == 0 || has1)) {
      int* dst = a.boxes + (col0 + col) * 4 + 2 * h;
      dst[0] = b0;
      dst[1] = b1;
    }
  }
  __syncthreads();
  // cut + re-spline (:396-401): the 8 one-sided eliminations (4 sequences x 2 sides, at most WARM = 32 rows each) run one per warp,
  // one row per lane: a row's step d -> (r - lo d) w is the affine map d -> a d + b with a = -lo w, b = r w, and the maps are
  // composed in row order by a shuffle tree (5 rounds) instead of a 32-step chain on one thread.  Then the 2x2 solves on 4 threads.
  {
    static_assert(G::WARM == 32, "one elimination row per lane");
    const int e = t >> 5, lane = t & 31, q = e >> 1, side = e & 1;
    const int b0 = sm.box[2 * q], b1 = sm.box[2 * q + 1];
    double fa = 1., fb = 0.;                             // identity for lanes without a row
    if (wallish_gap_ok(b0, b1)) {
      const int i = wallish_gap_row(b0, b1, side, lane);
      if (i >= 0) {
        const bool edge = (i == 0 || i == G::H - 1);
        const double w = wpivot(sm.wtab, side ? G::H - 1 - i : i, G::H);
        fa = edge ? 0. : -w;
        fb = wallish_gap_rhs(sm.B, q >> 1, q & 1, i) * w;
      }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {             // lane l ends up with the composition of rows l .. l + 2 off - 1, in order
      const double ha = __shfl_down_sync(0xffffffffu, fa, off), hb = __shfl_down_sync(0xffffffffu, fb, off);
      if (lane + off < 32) { fb = fma(ha, fb, hb); fa = ha * fa; }
    }
    if (lane == 0) sm.red[e] = fb;                       // reduced right-hand side at the last eliminated row (zero inflow)
  }
  __syncthreads();
  if (t < 4) sm.gaps[t] = wallish_gap_finish(sm.B, t >> 1, t & 1, sm.box[2 * t], sm.box[2 * t + 1], sm.red[2 * t], sm.red[2 * t + 1], sm.wtab);
  __syncthreads();
  {                                                                                         // :402
    // only the knots of the removed box change (64 threads per sequence); a box that reaches the end of the array
    // leaves NaN from b0 on, as the reference's spline does beyond its last knot
    const int q = t >> 6, h = q >> 1, col = q & 1;
    const WallishGap g = sm.gaps[q];
    const int iend = g.ok ? g.b1 : G::H - 1;
    for (int i = g.b0 + (t & 63); i <= iend; i += 64) {
      double* y = reinterpret_cast<double*>(sm.B + wpos(h, i)) + col;
      *y = wallish_fill(*y, i, g);
    }
  }
  __syncthreads();
  // DST-III (:409-412)
  wallish_dst3_pre(t, sm.B, v, a.twd);
  __syncthreads();                                      // the coefficients are in registers: the buffer is the FFT's again
}

// knots of the spliced spline (:413-419): exp(.)/k of the DST-III output on the kept rows, the unfiltered spectrum on the edge knots
// ROLL: raw values to their slots, exp(.)/k applied in place by a rolled loop
template <bool ROLL>
__device__ __forceinline__ void wallish_knots(const WallishArgs& a, const WallishSmem& sm, const int t, const long long col0, const bool has1,
                                              const double2 (&v)[16]) {
  typedef WallishGeo G;
  const double inv = 1. / G::N;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    double sign;
    const int j = wallish_dst3_out_index(t + 256 * r, sign);
    if (j >= a.i0 && j < a.i1) {
      if (ROLL) sm.B[ypos(a.lz + (j - a.i0))] = v[r];
      else {
        const double rk = __ldg(a.rklin + j);
        sm.B[ypos(a.lz + (j - a.i0))] = mk2(fast_exp(sign * inv * v[r].x) * rk, fast_exp(sign * inv * v[r].y) * rk);
      }
    }
  }
  for (int c = t; c < a.lz + a.rz; c += G::T) {
    const int row = c < a.lz ? a.nl - a.lz + c : a.nk - a.nr + (c - a.lz);
    const int knot = c < a.lz ? c : a.nmid + c;
    sm.B[ypos(knot)] = load_pair(a.pkout, a.ld, row, col0, has1, a.vec);
  }
  if (ROLL) {
    __syncthreads();
#pragma unroll 1
    for (int c0 = t; c0 < a.nmid; c0 += 4 * G::T) {       // four knots per step: their 1/k and shared-memory loads are in flight together
      double rk[4];
      double2 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + G::T * i;
        if (c < a.nmid) { rk[i] = __ldg(a.rklin + a.i0 + c); x[i] = sm.B[ypos(a.lz + c)]; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + G::T * i;
        if (c < a.nmid) {
          const double sc = ((a.i0 + c) & 1) ? -inv : inv;   // sign of wallish_dst3_out_index: odd rows -1, even rows +1
          sm.B[ypos(a.lz + c)] = mk2(fast_exp(sc * x[i].x) * rk[i], fast_exp(sc * x[i].y) * rk[i]);
        }
      }
    }
  }
  __syncthreads();
}

// clamped spline on the spliced knots (:420), chunks in registers; evaluation at self.k and blend (:420-423)
__device__ __forceinline__ void wallish_final(const WallishArgs& a, const WallishSmem& sm, const int t, const long long col0, const bool has1,
                                              double2 (&d)[16]) {
  typedef WallishGeo G;
  wallish_fin_forward_local(t, a.nc, sm.B, a.fc, d, sm.E, sm.Mf);
  __syncthreads();
  wallish_fin_fix_backward_local(t, a.fc, d, sm.E, sm.Mf, sm.Eb, sm.Mb, WallishCtaSync());
  __syncthreads();
  wallish_fin_backward_fix(t, a.fc, d, sm.Eb, sm.Mb);
  // the slopes next to an output wavenumber travel through the slot array
  double2* SL = sm.B + a.slbase;
  for (int round = 0; round < a.nrounds; ++round) {
    if (round) __syncthreads();
    wallish_fin_scatter(t, round, a.slotT, d, SL);
    __syncthreads();
    int q0 = __ldg(a.qstart + round), q1 = __ldg(a.qstart + round + 1);
    if (q0 < a.nl) q0 = a.nl;                           // the knots themselves were copied by wallish_copy_edges
    if (q1 > a.nk - a.nr) q1 = a.nk - a.nr;
    for (int q = q0 + t; q < q1; q += 2 * G::T) {       // two output wavenumbers per step: their table and spectrum loads overlap
      const int qb = q + G::T;
      const bool two = qb < q1;
      const double2 pka = load_pair(a.pkout, a.ld, q, col0, has1, a.vec);
      const double2 pkb = two ? load_pair(a.pkout, a.ld, qb, col0, has1, a.vec) : pka;
      const double2 ra = wallish_fin_eval(q, a.qinfo, a.qh, sm.B, SL, pka, __ldg(a.qth + q));
      const double2 rb = wallish_fin_eval(two ? qb : q, a.qinfo, a.qh, sm.B, SL, pkb, __ldg(a.qth + (two ? qb : q)));
      store_pair(a.pknow, a.ld, q, col0, has1, a.vec, ra);
      if (two) store_pair(a.pknow, a.ld, qb, col0, has1, a.vec, rb);
    }
  }
}

// V (lab): bit 0: the two FFTs share their code (a two-trip loop); bit 1: rolled logarithms; bit 2: rolled exponentials
// Two CTAs per SM: the 128-register budget of the register FFT.  The shared memory would admit three (kWallishSmemBytes), but the
// 80-register build that goes with them spills in every phase: 9.4 M P(k)/s against 11.6 M (profiles/r2_experiments.md, r3l).
template <int V>
__global__ void __launch_bounds__(256, 2) wallish_fused_kernel(const WallishArgs a) {
  extern __shared__ double2 smem_raw[];
  const WallishSmem sm(smem_raw);
  const int t = threadIdx.x;
  const long long npairs = (a.ncols + 1) / 2;
  __shared__ __align__(8) uint64_t s_bar;        // completion of the bulk copies of the rows layout
  unsigned bar_parity = 0;
  if (t == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  if (t < 32) sm.wtab[t] = a.wtab[t];
  __syncthreads();
  for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const long long col0 = 2 * pair;
    const bool has1 = col0 + 1 < a.ncols;     // odd column count: the last pair repeats its column
    double2 v[16];
    double2 d[16];
    long long c0 = clock64();
    auto stamp = [&](const int ph) {
      if (a.dbg && t == 0) { const long long c1 = clock64(); atomicAdd(a.dbg + ph, (unsigned long long)(c1 - c0)); c0 = c1; }
    };
    if (V & 1) {
#pragma unroll 1
      for (int leg = 0; leg < 2; ++leg) {
        if (leg == 0) wallish_load_log<(V & 2) != 0>(a, sm, t, col0, has1, v, &s_bar, bar_parity);
        fft4096(t, v, sm.B, a.tw1, a.tw2);
        if (leg == 0) wallish_middle(a, sm, t, col0, has1, v, d);
      }
    } else {
      wallish_load_log<(V & 2) != 0>(a, sm, t, col0, has1, v, &s_bar, bar_parity);
      stamp(0);
      fft4096(t, v, sm.B, a.tw1, a.tw2);
      stamp(1);
      wallish_middle(a, sm, t, col0, has1, v, d);
      stamp(2);
      fft4096(t, v, sm.B, a.tw1, a.tw2);
    }
    __syncthreads();
    stamp(3);
    wallish_knots<(V & 4) != 0>(a, sm, t, col0, has1, v);
    stamp(4);
    wallish_final(a, sm, t, col0, has1, d);
    stamp(5);
    fence_proxy_async_smem();                             // this pair's writes to the buffer are ordered before the next pair's bulk copies into it
    __syncthreads();                                      // the buffer goes back to the next pair's FFT
  }
}

// ---- standalone DST-II / DST-III (orthonormal), N = 4096, along axis 0 of [4096, ncols] ------------------------------
template <int TYPE>
__global__ void __launch_bounds__(256, 2) dst_kernel(const double* __restrict__ in, double* __restrict__ out, const long long ncols,
                                                     const double2* tw1, const double2* tw2, const double2* twd) {
  typedef WallishGeo G;
  extern __shared__ double2 smem_raw[];
  double2* B = smem_raw;
  const int t = threadIdx.x;
  const long long col0 = 2LL * blockIdx.x;
  const bool has1 = col0 + 1 < ncols;
  double2 v[16];
  if (TYPE == 2) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = makhoul_src(t + 256 * r, sign);
      const double* row = in + (long long)j * ncols + col0;
      v[r] = mk2(sign * row[0], has1 ? sign * row[1] : 0.);
    }
    dst2_in_smem(t, v, B, tw1, tw2, twd);
    for (int kk = t; kk < G::N; kk += 256) {
      const double2 x = B[wpos(kk & 1, kk >> 1)];
      double* dst = out + (long long)kk * ncols + col0;
      dst[0] = x.x;
      if (has1) dst[1] = x.y;
    }
  } else {
    for (int kk = t; kk < G::N; kk += 256) {
      const double* row = in + (long long)kk * ncols + col0;
      B[wpos(kk & 1, kk >> 1)] = mk2(row[0], has1 ? row[1] : 0.);
    }
    __syncthreads();
    wallish_dst3_pre(t, B, v, twd);
    __syncthreads();
    fft4096(t, v, B, tw1, tw2);
    const double inv = 1. / G::N;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = wallish_dst3_out_index(t + 256 * r, sign);
      double* dst = out + (long long)j * ncols + col0;
      dst[0] = sign * inv * v[r].x;
      if (has1) dst[1] = sign * inv * v[r].y;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
int upload(void** dptr, const void* src, size_t bytes);   // cpf_fftlog.cu

static std::mutex g_wt_mutex;
static std::vector<WallishTables> g_wt;

static int wallish_tables(int device, WallishTables* out) {
  std::lock_guard<std::mutex> lock(g_wt_mutex);
  for (auto& e : g_wt)
    if (e.device == device) { *out = e; return CPF_OK; }
  const int N = WallishGeo::N;
  const long double PI = acosl(-1.0L);
  static const int expo[6] = {1, 2, 3, 4, 8, 12};
  std::vector<double2> tw1(6 * 256), tw2(6 * 16), twd(N);
  auto root = [&](long long num, long long den) {
    const long double ang = -2.0L * PI * (long double)(num % den) / (long double)den;
    double2 r;
    r.x = (double)cosl(ang);
    r.y = (double)sinl(ang);
    return r;
  };
  for (int e = 0; e < 6; ++e) {
    for (int n2 = 0; n2 < 256; ++n2) tw1[e * 256 + n2] = root((long long)expo[e] * n2, N);
    for (int m2 = 0; m2 < 16; ++m2) tw2[e * 16 + m2] = root(expo[e] * m2, 256);
  }
  for (int k = 0; k < N; ++k) twd[k] = root(k, 4LL * N);   // exp(-i pi k / 2N) = exp(-2 pi i k / 4N)
  double wtab[32];
  wtab[0] = 1.;
  double c = 0.;
  for (int i = 1; i < 32; ++i) { wtab[i] = 1. / (4. - c); c = wtab[i]; }
  WallishTables wt;
  wt.device = device;
  CPF_TRY(upload((void**)&wt.tw1, tw1.data(), tw1.size() * sizeof(double2)));
  CPF_TRY(upload((void**)&wt.tw2, tw2.data(), tw2.size() * sizeof(double2)));
  CPF_TRY(upload((void**)&wt.twd, twd.data(), twd.size() * sizeof(double2)));
  CPF_TRY(upload((void**)&wt.wtab, wtab, sizeof(wtab)));
  g_wt.push_back(wt);
  *out = wt;
  return CPF_OK;
}

// Everything cpf_wallish2018 derives from the two wavenumber grids (bao_filter.py:364, 413-431), cached per (device, klin, kout) together
// with its device image: [facT | qh | qth | 1/klin | klin] (doubles) then [slotT | qstart | qinfo] (ints).  Entries live until evicted
// (at most kWallishPlans per process; the eviction synchronises the device before it frees the image).
struct WallishGridPlan {
  int device, cap_limit;
  std::vector<double> klin, kout;
  WallishFinalPlan fp;
  void* d_tab = nullptr;
  const double* d_klin() const { return reinterpret_cast<const double*>(d_tab) + fp.facT.size() + fp.qh.size() + kout.size() + klin.size(); }
};
static std::mutex g_gp_mutex;
static std::vector<WallishGridPlan*> g_gp;
constexpr size_t kWallishPlans = 8;

static int wallish_grid_plan(int device, const std::vector<double>& klin, const std::vector<double>& kout, int cap_limit, const WallishGridPlan** out) {
  std::lock_guard<std::mutex> lock(g_gp_mutex);
  for (size_t i = 0; i < g_gp.size(); ++i) {
    WallishGridPlan* e = g_gp[i];
    if (e->device == device && e->cap_limit == cap_limit && e->klin.size() == klin.size() && e->kout.size() == kout.size() &&
        memcmp(e->klin.data(), klin.data(), klin.size() * sizeof(double)) == 0 && memcmp(e->kout.data(), kout.data(), kout.size() * sizeof(double)) == 0) {
      g_gp.erase(g_gp.begin() + i);          // most recently used last
      g_gp.push_back(e);
      *out = e;
      return CPF_OK;
    }
  }
  const int nlin = (int)klin.size(), nk = (int)kout.size();
  for (int i = 1; i < nlin; ++i) if (!(klin[i] > klin[i - 1])) return fail(CPF_EINVAL, "cpf_wallish2018: klin must be strictly increasing");
  for (int i = 1; i < nk; ++i) if (!(kout[i] > kout[i - 1])) return fail(CPF_EINVAL, "cpf_wallish2018: kout must be strictly increasing");
  WallishGridPlan* e = new WallishGridPlan();
  e->device = device; e->cap_limit = cap_limit; e->klin = klin; e->kout = kout;
  const std::string why = wallish_final_plan(klin.data(), nlin, kout.data(), nk, &e->fp, cap_limit);
  if (!why.empty()) { delete e; return fail(CPF_EINVAL, "cpf_wallish2018: %s", why.c_str()); }
  const WallishFinalPlan& fp = e->fp;
  const size_t n_fac = fp.facT.size(), n_qh = fp.qh.size(), n_slot = fp.slotT.size(), n_qs = fp.qstart.size(), n_qi = fp.qinfo.size();
  const size_t tab_bytes = (n_fac + n_qh + (size_t)nk + 2 * (size_t)nlin) * sizeof(double) + (n_slot + n_qs + n_qi) * sizeof(int);
  std::vector<char> h_tab(tab_bytes);
  double* pd = reinterpret_cast<double*>(h_tab.data());
  memcpy(pd, fp.facT.data(), n_fac * sizeof(double));
  memcpy(pd + n_fac, fp.qh.data(), n_qh * sizeof(double));
  for (int q = 0; q < nk; ++q) {                                          // _tophat(self.k, kmax=1, scale=20) (:425-431)
    const double k = kout[q];
    pd[n_fac + n_qh + q] = k > 1. ? exp(-400. * (k - 1.) * (k - 1.)) : 1.;
  }
  for (int i = 0; i < nlin; ++i) { pd[n_fac + n_qh + nk + i] = 1. / klin[i]; pd[n_fac + n_qh + nk + nlin + i] = klin[i]; }
  int* pi = reinterpret_cast<int*>(pd + n_fac + n_qh + nk + 2 * (size_t)nlin);
  memcpy(pi, fp.slotT.data(), n_slot * sizeof(int));
  memcpy(pi + n_slot, fp.qstart.data(), n_qs * sizeof(int));
  memcpy(pi + n_slot + n_qs, fp.qinfo.data(), n_qi * sizeof(int));
  if (int rc = upload(&e->d_tab, h_tab.data(), tab_bytes)) { delete e; return rc; }
  if (g_gp.size() >= kWallishPlans) {
    WallishGridPlan* old = g_gp.front();
    g_gp.erase(g_gp.begin());
    DeviceGuard g(old->device);
    cudaDeviceSynchronize();               // a launch may still read the image
    cudaFree(old->d_tab);
    delete old;
  }
  g_gp.push_back(e);
  *out = e;
  return CPF_OK;
}

static int check_device(const char* who, int device) {
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "%s: device %d out of range (%d visible)", who, device, ndev);
  return CPF_OK;
}

}  // namespace cpf

using namespace cpf;

extern "C" {

int cpf_dst(int type, const double* in, int nx, int64_t ncols, double* out, int on_device, int device, void* stream_) {
  if (type != 2 && type != 3) return fail(CPF_EINVAL, "cpf_dst: type must be 2 or 3, got %d", type);
  if (nx != WallishGeo::N) return fail(CPF_EUNSUPPORTED, "cpf_dst: only nx = %d (the Wallish2018 grid) is implemented, got %d", WallishGeo::N, nx);
  if (ncols < 0) return fail(CPF_EINVAL, "cpf_dst: negative column count");
  if (ncols == 0) return CPF_OK;
  if (!in || !out) return fail(CPF_EINVAL, "cpf_dst: null buffer");
  CPF_TRY(check_device("cpf_dst", device));
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  WallishTables wt;
  CPF_TRY(wallish_tables(device, &wt));
  const size_t bytes = (size_t)nx * (size_t)ncols * sizeof(double);
  ScratchBuf din, dout;
  const double* d_in = in;
  double* d_out = out;
  if (!on_device) {
    CPF_CUDA(din.alloc(bytes, stream));
    CPF_CUDA(dout.alloc(bytes, stream));
    CPF_CUDA(cudaMemcpyAsync(din.p, in, bytes, cudaMemcpyHostToDevice, stream));
    d_in = (const double*)din.p;
    d_out = (double*)dout.p;
  }
  const unsigned grid = (unsigned)((ncols + 1) / 2);
  if (type == 2) {
    CPF_CUDA(cudaFuncSetAttribute(dst_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWallishSmemBytes));
    dst_kernel<2><<<grid, 256, kWallishSmemBytes, stream>>>(d_in, d_out, ncols, wt.tw1, wt.tw2, wt.twd);
  } else {
    CPF_CUDA(cudaFuncSetAttribute(dst_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWallishSmemBytes));
    dst_kernel<3><<<grid, 256, kWallishSmemBytes, stream>>>(d_in, d_out, ncols, wt.tw1, wt.tw2, wt.twd);
  }
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  return CPF_OK;
}

static int wallish2018_impl(const double* klin, const double* pklin, int nlin, const double* kout, const double* pkout, int nk,
                            int64_t ncols, double* pknow, int32_t* boxes, int on_device, int device, void* stream_, const int lin_rows) {
  if (nlin != WallishGeo::N) return fail(CPF_EUNSUPPORTED, "cpf_wallish2018: nlin must be %d (bao_filter.py:364), got %d", WallishGeo::N, nlin);
  if (nk < 2) return fail(CPF_EINVAL, "cpf_wallish2018: nk = %d", nk);
  if (ncols < 0) return fail(CPF_EINVAL, "cpf_wallish2018: negative column count");
  if (ncols == 0) return CPF_OK;
  if (!klin || !pklin || !kout || !pkout || !pknow) return fail(CPF_EINVAL, "cpf_wallish2018: null buffer");
  CPF_TRY(check_device("cpf_wallish2018", device));
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  WallishTables wt;
  CPF_TRY(wallish_tables(device, &wt));

  // The two wavenumber grids are needed on the host (knot selection, interval search, spline factors); everything derived from them is cached
  // per (device, klin, kout) with its device copy (wallish_grid_plan).  The rows entry takes the grids as HOST arrays: no device-to-host copy,
  // no synchronisation, and with device spectra the call is asynchronous on the stream.
  std::vector<double> h_klin(nlin), h_kout(nk);
  if (on_device && !lin_rows) {
    CPF_CUDA(cudaMemcpyAsync(h_klin.data(), klin, nlin * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaMemcpyAsync(h_kout.data(), kout, nk * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  } else {
    h_klin.assign(klin, klin + nlin);
    h_kout.assign(kout, kout + nk);
  }
  const char* cap_env = getenv("CPF_WALLISH_SLOT_CAP");                   // tests: forces the multi-round evaluation
  const WallishGridPlan* gp = nullptr;
  CPF_TRY(wallish_grid_plan(device, h_klin, h_kout, cap_env ? atoi(cap_env) : 0, &gp));
  const WallishFinalPlan& fp = gp->fp;

  const size_t lin_bytes = (size_t)nlin * ncols * sizeof(double), out_bytes = (size_t)nk * ncols * sizeof(double);
  ScratchBuf d_pklin, d_pkout, d_pknow, d_boxes;
  const double *p_klin = gp->d_klin(), *p_pklin = pklin, *p_pkout = pkout;
  double* p_pknow = pknow;
  int* p_boxes = boxes;
  if (!on_device) {
    CPF_CUDA(d_pklin.alloc(lin_bytes, stream));
    CPF_CUDA(d_pkout.alloc(out_bytes, stream));
    CPF_CUDA(d_pknow.alloc(out_bytes, stream));
    CPF_CUDA(cudaMemcpyAsync(d_pklin.p, pklin, lin_bytes, cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(d_pkout.p, pkout, out_bytes, cudaMemcpyHostToDevice, stream));
    p_pklin = (const double*)d_pklin.p; p_pkout = (const double*)d_pkout.p;
    p_pknow = (double*)d_pknow.p;
    if (boxes) {
      CPF_CUDA(d_boxes.alloc((size_t)ncols * 4 * sizeof(int), stream));
      p_boxes = (int*)d_boxes.p;
    }
  }
  WallishArgs a;
  a.klin = p_klin; a.pklin = p_pklin; a.pkout = p_pkout; a.pknow = p_pknow; a.ncols = ncols; a.ld = ncols;
  a.lin_rows = lin_rows ? 1 : 0;
  if (lin_rows && (uintptr_t)p_pklin % 16 != 0) return fail(CPF_EINVAL, "cpf_wallish2018_rows: pklin must be 16-byte aligned");
  a.vec = (ncols % 2 == 0) && ((uintptr_t)p_pklin % 16 == 0) && ((uintptr_t)p_pkout % 16 == 0) && ((uintptr_t)p_pknow % 16 == 0);
  a.boxes = p_boxes;
  a.dbg = nullptr;
  ScratchBuf d_dbg;
  if (getenv("CPF_WALLISH_DBG")) { CPF_CUDA(d_dbg.alloc(8 * sizeof(unsigned long long), stream)); CPF_CUDA(cudaMemsetAsync(d_dbg.p, 0, 64, stream)); a.dbg = (unsigned long long*)d_dbg.p; }
  a.tw1 = wt.tw1; a.tw2 = wt.tw2; a.twd = wt.twd; a.wtab = wt.wtab;
  a.i0 = fp.i0; a.i1 = fp.i1; a.nl = fp.nl; a.nr = fp.nr; a.lz = fp.lz; a.rz = fp.rz; a.nmid = fp.nmid; a.nc = fp.nc; a.nk = nk;
  a.nrounds = fp.nrounds; a.slbase = fp.slbase;
  a.fc.facT = reinterpret_cast<const double*>(gp->d_tab);
  a.fc.t0 = fp.ut0; a.fc.t1 = fp.ut1; a.fc.Lw = fp.uLw; a.fc.cp = fp.ucp; a.fc.P = fp.uP; a.fc.Q = fp.uQ;
  a.qh = a.fc.facT + fp.facT.size();
  a.qth = a.qh + fp.qh.size();
  a.rklin = a.qth + nk;
  a.slotT = reinterpret_cast<const int*>(a.rklin + 2 * (size_t)nlin);        // (the device copy of klin sits behind rklin)
  a.qstart = a.slotT + fp.slotT.size();
  a.qinfo = a.qstart + fp.qstart.size();
  int sms = 0;
  CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const long long npairs = (ncols + 1) / 2;
  unsigned grid = (unsigned)npairs;
  int variant = 2;
  if (const char* e = getenv("CPF_WALLISH_PERSISTENT")) { if (e[0] == '1' && npairs > 2LL * sms) grid = 2u * (unsigned)sms; }
  if (const char* e = getenv("CPF_WALLISH_VARIANT")) variant = atoi(e) & 7;
  typedef void (*wkern_t)(const WallishArgs);
  static const wkern_t kerns[8] = {wallish_fused_kernel<0>, wallish_fused_kernel<1>, wallish_fused_kernel<2>, wallish_fused_kernel<3>,
                                   wallish_fused_kernel<4>, wallish_fused_kernel<5>, wallish_fused_kernel<6>, wallish_fused_kernel<7>};
  wkern_t kern = kerns[variant];
  size_t smem_bytes = kWallishSmemBytes;
  if (const char* e = getenv("CPF_WALLISH_SMEM_PAD")) smem_bytes += (size_t)atoi(e);          // lab: lowers the occupancy
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  if (a.dbg) {
    int occ = 0;
    CPF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem_bytes));
    fprintf(stderr, "wallish: %d resident CTAs per SM\n", occ);
  }
  kern<<<grid, 256, smem_bytes, stream>>>(a);
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(pknow, p_pknow, out_bytes, cudaMemcpyDeviceToHost, stream));
    if (boxes) CPF_CUDA(cudaMemcpyAsync(boxes, p_boxes, (size_t)ncols * 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  }
  // host results must have landed when the call returns; device results are stream-ordered (scratch memory too)
  if (!on_device || a.dbg) CPF_CUDA(cudaStreamSynchronize(stream));
  if (a.dbg) {
    unsigned long long h[8];
    CPF_CUDA(cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "wallish phases (cycles per pair, thread 0): load+log %.0f  fft %.0f  middle %.0f  fft %.0f  knots %.0f  final %.0f\n", (double)h[0] / npairs,
            (double)h[1] / npairs, (double)h[2] / npairs, (double)h[3] / npairs, (double)h[4] / npairs, (double)h[5] / npairs);
  }
  return CPF_OK;
}

int cpf_wallish2018(const double* klin, const double* pklin, int nlin, const double* kout, const double* pkout, int nk,
                    int64_t ncols, double* pknow, int32_t* boxes, int on_device, int device, void* stream) {
  return wallish2018_impl(klin, pklin, nlin, kout, pkout, nk, ncols, pknow, boxes, on_device, device, stream, 0);
}

int cpf_wallish2018_rows(const double* klin, const double* pklin_rows, int nlin, const double* kout, const double* pkout, int nk,
                         int64_t ncols, double* pknow, int32_t* boxes, int on_device, int device, void* stream) {
  return wallish2018_impl(klin, pklin_rows, nlin, kout, pkout, nk, ncols, pknow, boxes, on_device, device, stream, 1);
}

}  // extern "C"
