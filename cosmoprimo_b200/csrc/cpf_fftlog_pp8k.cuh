// cpf_fftlog_pp8k.cuh — persistent FFTLog kernel (sm_100a) for N = 8192 (nk = 4096: cosmoprimo/fftlog.py:149-150), default call
// (zero padding, cropped output, real post-factor: fftlog.py:198-241).  Included by cpf_fftlog.cu after cpf_fftlog_pp.cuh.
//
// The 8192-point transforms are split by one radix-2 stage into two 4096-point "chains" (as in fftlog_split2_kernel, which stays the
// kernel of small launches and of every other option):
//   FFT #1, decimation in frequency: bins 2k + c = FFT_4096( a[n] w_8192^{c n} )[k]      (the upper half of the rotated input is zero padding)
//   FFT #2, decimation in time:      g[j] = E[j] + w_8192^j O[j], j < 4096 (cropped),      E / O = FFT_4096 of the even / odd bins times u
// One 512-thread CTA per SM stays resident; its two 256-thread groups are the two chains of ONE pair of rows at a time:
//   * the two rows of the CTA's next pair arrive by TMA bulk copies (2 x 32 KB) one pair ahead and are read by BOTH groups: every sample
//     crosses L2 -> SM once (the split kernel loads each row twice, from global memory, at the top of every CTA);
//   * the thread-private tables -- 16 pass-1 twiddles, the 16 + 16 kernel-spectrum values of the two chains, 16 + 16 pre/post factors --
//     live in tensor memory (all 512 columns; tcgen05.ld, double buffered), the pass-2 twiddles (16 x 16, they depend on t % 16 only) in
//     shared memory with conflict-free 128-bit reads; only the chain-1 twiddles w_8192^{t + 256 r} come from global memory (L1/L2 hits);
//   * each chain is the ping-pong kernel's register FFT (cpf_fft_core.h) on the group's own exchange buffer with group barriers; the
//     groups meet three times per pair: after the input is in registers (the staging buffer is refilled), and around the exchange of
//     the chain results (group 0 finishes output rows 0..7, group 1 rows 8..15);
//   * the chains run the same phases, so chain 1 is held back by 0.8 us after the first of those meetings: its shared-memory phases then fall
//     on chain 0's fp64 phases (+2.5 %);
//   * programmatic dependent launch and non-finite handling as in the other persistent kernels.
#pragma once

#include "cpf_fft_core.h"

namespace cpf {

// per-thread table record [P, 256, P8_REC] (build_pp8k_tables): [0,16) pass-1 twiddles w_4096^{t k1} ; [16,32) / [32,48) kernel spectrum of
// chain 0 / 1 at the 8192-point bins 2 (t + 256 r) + c (Hermitian-extended, 1/N and the (-1)^k of the N/4 rotation folded in) ;
// [48,56) pre [N/4 + t + 256 r], r < 16 (16 doubles) ; [56,64) post, the same
constexpr int P8_REC = 64;
constexpr uint32_t P8_COL_TW1 = 0, P8_COL_UT0 = 64, P8_COL_UT1 = 128, P8_COL_PRE = 192, P8_COL_POST = 224;
constexpr int P8_SMEM_ELEMS = 2 * Geo<16>::SMEM_ELEMS + 256;
constexpr int P8_SMEM_BYTES = P8_SMEM_ELEMS * (int)sizeof(double2);
constexpr int P8_SMEM_BYTES_TMA = P8_SMEM_BYTES + 2 * 4096 * (int)sizeof(double);

// w_32^r = exp(-2 pi i r / 32), r < 16
__device__ constexpr double kW32[16][2] = {{1.00000000000000000000e+00, -0.00000000000000000000e+00}, {9.80785280403230430579e-01, -1.95090322016128248084e-01}, {9.23879532511286738483e-01, -3.82683432365089781779e-01}, {8.31469612302545235671e-01, -5.55570233019602177649e-01}, {7.07106781186547572737e-01, -7.07106781186547461715e-01}, {5.55570233019602288671e-01, -8.31469612302545235671e-01}, {3.82683432365089837290e-01, -9.23879532511286738483e-01}, {1.95090322016128331351e-01, -9.80785280403230430579e-01}, {6.12323399573676603587e-17, -1.00000000000000000000e+00}, {-1.95090322016128192573e-01, -9.80785280403230430579e-01}, {-3.82683432365089726268e-01, -9.23879532511286738483e-01}, {-5.55570233019601955604e-01, -8.31469612302545457716e-01}, {-7.07106781186547461715e-01, -7.07106781186547572737e-01}, {-8.31469612302545346694e-01, -5.55570233019602177649e-01}, {-9.23879532511286738483e-01, -3.82683432365089892802e-01}, {-9.80785280403230430579e-01, -1.95090322016128608906e-01}};

template <bool TMA>
__global__ void __launch_bounds__(512, 1) fftlog_pp8k_kernel(const FftlogArgs a, const double2* __restrict__ tmtab, const double2* __restrict__ tw2tab,
                                                             const double2* __restrict__ tw8192) {
  typedef Geo<16> G;
  constexpr int T = 256, M = 4096, N = 8192, RS = G::RS;
  extern __shared__ double2 smem[];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(8) uint64_t s_mbar;
  const int warp = threadIdx.x >> 5;
  const int g = threadIdx.x >> 8, t = threadIdx.x & 255;
  double2* S = smem + g * G::SMEM_ELEMS;
  const double2* Sother = smem + (1 - g) * G::SMEM_ELEMS;
  double2* TW2 = smem + 2 * G::SMEM_ELEMS;                                  // [l1][m2] = w_256^{m2 l1}
  double* stage = reinterpret_cast<double*>(smem + P8_SMEM_ELEMS);          // TMA only: row a [4096], row b [4096]
  unsigned stage_parity = 0;
  if (TMA && threadIdx.x == 0) { mbar_init(&s_mbar, 1); mbar_fence_init(); }
  auto stage_rows = [&](const int p, const long long pair) {              // one thread per CTA
    const long long b0 = 2 * pair;
    const bool two = b0 + 1 < a.batch;
    mbar_expect_tx(&s_mbar, two ? 2u * M * 8u : M * 8u);
    bulk_g2s(stage, a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n, M * 8u, &s_mbar);
    if (two) bulk_g2s(stage + M, a.in + (a.in_has_P ? ((b0 + 1) * a.P + p) : b0 + 1) * (long long)a.n, M * 8u, &s_mbar);
  };

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) tmem_alloc_all(&s_tmem_base);
  if (threadIdx.x < 256) TW2[threadIdx.x] = tw2tab[threadIdx.x];
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  // lane quarter of this warp, column half of this thread (t / 128); both groups read the same copy
  const uint32_t tb = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u * (uint32_t)(t >> 7);
  auto gbar = [&]() { named_sync(1 + g, T); };

  // DFT outputs w[0..16) times the thread's 16 twiddles in tensor memory at column `col` (chunks of 4, the next chunk in flight while
  // this one is used), scattered to dst[k * stride]; does the DFT of w first
  auto twiddle_store_tm = [&](double2 (&w)[16], const uint32_t col, double2* dst, const int stride) {
    Tm4 tw[2];
    tmem_ld4(col, tw[0]);
    dft_dit<16, false, false>(w);             // the first chunk of twiddles arrives behind the butterflies
    tmem_wait4(tw[0]);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < 3) tmem_ld4(col + 16 * (ch + 1), tw[(ch + 1) & 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * ch + i;
        if (k > 0) w[k] = cmul(w[k], tw[ch & 1].get(i));
      }
      if (ch < 3) tmem_wait4(tw[(ch + 1) & 1]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k * stride] = w[k];
  };
  auto pass1 = [&](const double2 (&v)[16]) {
    double2 w[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) w[bitrev(n1, 4)] = v[n1];
    twiddle_store_tm(w, tb + P8_COL_TW1, S + t, RS);
  };
  auto pass2 = [&]() {
    const int k1 = t >> 4, m2 = t & 15;
    double2* row = S + k1 * RS + m2;
    double2 w[16];
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) w[bitrev(m1, 4)] = row[16 * m1];
    dft_dit<16, false, false>(w);
    row[0] = w[0];
#pragma unroll
    for (int l1 = 1; l1 < 16; ++l1) row[16 * l1] = cmul(w[l1], TW2[16 * l1 + m2]);
  };
  auto pass3_load = [&](double2 (&v)[16]) {
    const int k1 = t & 15, l1 = t >> 4;
    const double2* row = S + k1 * RS + 16 * l1;
#pragma unroll
    for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = row[m2];
  };
  // chain 1: times w_8192^{t + 256 r} = w_8192^t w_32^r (input of FFT #1, output of FFT #2): one 16-byte load and 15 products with
  // compile-time constants instead of 16 loads (the loads were 1 k of the 17 k LSU wavefronts per pair and missed L1)
  auto chain_twiddle = [&](double2 (&v)[16]) {
    if (g == 1) {
      const double2 base = __ldg(tw8192 + t);
      v[0] = cmul(v[0], base);
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], cmul(base, mk2(kW32[r][0], kW32[r][1])));
    }
  };

  // Work split: with at least as many CTAs as plan rows every CTA serves ONE plan row (its tables are loaded once) and the CTAs of a row
  // take its pairs in turn; otherwise every CTA walks over all plan rows.
  const bool one_row = (int)gridDim.x >= a.P;
  const int p_lo = one_row ? (int)blockIdx.x % a.P : 0, p_hi = one_row ? p_lo + 1 : a.P;
  const long long first = one_row ? (long long)blockIdx.x / a.P : (long long)blockIdx.x;
  const long long stride = one_row ? ((long long)gridDim.x - p_lo + a.P - 1) / a.P : (long long)gridDim.x;
  for (int p = p_lo; p < p_hi; ++p) {
    if (p > p_lo) { tmem_fence_before(); __syncthreads(); tmem_fence_after(); }   // all reads of the old tables are done
    if (threadIdx.x < 256) {
      const double2* rec = tmtab + ((size_t)p * T + t) * P8_REC;
#pragma unroll 2
      for (int ch = 0; ch < P8_REC / 4; ++ch) {
        double2 d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = rec[4 * ch + i];
        tmem_st4(tb + 16 * ch, d);
      }
      tmem_wait_st();
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    if (p == p_lo) asm volatile("griddepcontrol.wait;" ::: "memory");      // plan tables only so far; rows may come from the previous kernel
    if (TMA && threadIdx.x == 0 && first < a.pairs_per_p) stage_rows(p, first);

    for (long long pair = first; pair < a.pairs_per_p; pair += stride) {
      const long long b0 = 2 * pair, b1 = b0 + 1;
      const bool has1 = b1 < a.batch;
      const double* rowA = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
      const double* rowB = has1 ? a.in + (a.in_has_P ? (b1 * a.P + p) : b1) * (long long)a.n : rowA;
      if (!TMA) {                                                         // L2 prefetch of the CTA's next pair (one 128-byte line per thread and row)
        const long long nb0 = 2 * (pair + stride);
        if (nb0 < a.batch) {
          const int lines = (a.n * 8 + 127) / 128;
          for (int l = threadIdx.x; l < 2 * lines; l += 512) {
            const long long bn = nb0 + (l < lines ? 0 : 1);
            if (bn < a.batch) {
              const double* q = a.in + (a.in_has_P ? (bn * a.P + p) : bn) * (long long)a.n + (size_t)(l < lines ? l : l - lines) * 16;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
          }
        }
      }

      double2 v[16];
      bool bad_a = false, bad_b = false;
      // ---- the 4096 samples of the rotated window (the other 4096 are zero padding), times pre; both groups read the same rows ----
      if (TMA) { mbar_wait(&s_mbar, stage_parity); stage_parity ^= 1u; }
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        Tm4 tq;
        tmem_ld4(tb + P8_COL_PRE + 16 * hb, tq);
        double x[8], y[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int rr = 8 * hb + r;
          if (TMA) {
            x[r] = stage[t + T * rr];
            y[r] = has1 ? stage[M + t + T * rr] : 0.;
          } else {
            const int i = t + T * rr + N / 4 - a.in_left;
            const bool ok = (unsigned)i < (unsigned)a.n;
            x[r] = ok ? __ldcs(rowA + i) : 0.;
            y[r] = (ok && has1) ? __ldcs(rowB + i) : 0.;
          }
        }
        bool ba = false, bb = false;
        scrub_rows(x, y, ba, bb);
        bad_a |= ba; bad_b |= bb;
        tmem_wait4(tq);
#pragma unroll
        for (int r = 0; r < 8; ++r) { const double pr = tq.getd(r); v[8 * hb + r] = mk2(x[r] * pr, y[r] * pr); }
      }
      chain_twiddle(v);

      // ---- FFT #1 ----
      pass1(v);
      const bool row_a_bad = __syncthreads_or(bad_a);                     // every thread has its samples in registers
      if (a.ahead > 0 && g == 1) __nanosleep((unsigned)a.ahead);          // phase offset between the two chains (launch_pp8k)
      if (TMA && threadIdx.x == 0 && pair + stride < a.pairs_per_p) stage_rows(p, pair + stride);
      pass2();
      const bool row_b_bad = named_sync_or(1 + g, T, bad_b);              // both groups loaded the same rows: the same flags in both
      pass3_load(v);
      gbar();   // pass-3 reads of S are done before FFT #2 overwrites it
      // ---- kernel multiply: bins 2 (t + 256 r) + g ----
      {
        const uint32_t col = tb + (g ? P8_COL_UT1 : P8_COL_UT0);
        Tm4 tu[2];
        tmem_ld4(col, tu[0]);
        dft_dit<16, false, false>(v);
        tmem_wait4(tu[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (ch < 3) tmem_ld4(col + 16 * (ch + 1), tu[(ch + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[4 * ch + i] = cmul(v[4 * ch + i], tu[ch & 1].get(i));
          if (ch < 3) tmem_wait4(tu[(ch + 1) & 1]);
        }
      }
      // ---- FFT #2 ----
      pass1(v);
      gbar();
      pass2();
      gbar();
      pass3_load(v);
      gbar();
      dft_dit<16, false, false>(v);
      chain_twiddle(v);                                                   // O'[j] = w_8192^j O[j]
      // ---- g[j] = E[j] + O'[j]: group 0 finishes rows r < 8, group 1 rows r >= 8; the halves cross through the exchange buffers ----
      double2 res[8];
      if (g == 0) {                                                       // (register indices must be compile-time constants)
#pragma unroll
        for (int r = 0; r < 8; ++r) { S[t + T * r] = v[8 + r]; res[r] = v[r]; }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) { S[t + T * r] = v[r]; res[r] = v[8 + r]; }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const double2 w = Sother[t + T * r];
        res[r] = mk2(res[r].x + w.x, res[r].y + w.y);
      }
      Tm4 tq;
      tmem_ld4(tb + P8_COL_POST + 16 * g, tq);
      __syncthreads();                                                    // the other group has read this group's buffer: the next pair may scatter into it
      tmem_wait4(tq);
      double* outA = a.out + (size_t)(b0 * a.P + p) * a.n_out;
      double* outB = a.out + (size_t)(b1 * a.P + p) * a.n_out;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int o = t + T * (8 * g + r) + N / 4 - a.out_left;
        if ((unsigned)o < (unsigned)a.n_out) {
          const double po = tq.getd(r);
          __stcs(outA + o, row_a_bad ? nan("") : res[r].x * po);
          if (has1) __stcs(outB + o, row_b_bad ? nan("") : res[r].y * po);
        }
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_all(s_tmem_base);
}

}  // namespace cpf
