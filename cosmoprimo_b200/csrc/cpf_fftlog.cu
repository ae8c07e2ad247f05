// cpf_fftlog.cu — fused FFTLog kernels (sm_100a), plan object and the cpf_fftlog / cpf_rfft / cpf_irfft_conj entry
// points of include/cpfftlog.h.
//
// What the kernels compute (cosmoprimo/fftlog.py:198-241), for one input row f:
//     a = pad(f) * pre ;  A = rfft(a) ;  g = irfft(conj(A * u), n=N) ;  G = g * post ;  crop
// Reformulation used here:
//   * the map a -> g is REAL-linear: g = FFT_fwd( ut .* FFT_fwd(a) ) / N, where ut is u extended to N bins by
//     Hermitian symmetry with Im dropped at DC and Nyquist (numpy's irfft ignores them).  `conj` + inverse FFT is a
//     second forward FFT, so one forward routine serves both.
//   * two independent rows are packed as z = a + i b; Re/Im of the result are the two outputs.  No real-FFT
//     split/merge passes are needed at all.
//   * with zero extrapolation the non-zero input window and the cropped output window both sit inside
//     [N/4, 3N/4).  Rotating both by N/4 multiplies ut by (-1)^k and leaves an input whose upper half is zero and an
//     output whose upper half is not needed: the first radix stage of FFT#1 and the last of FFT#2 are pruned.
//
// One CTA handles one pair of rows (T = N/16 threads, 16 complex values per thread in registers, three register
// passes per FFT, two shared-memory exchanges per FFT); nothing but the input rows and the output rows touches HBM.
#include <cooperative_groups.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "cpf_common.h"
#include "cpf_fft_core.h"

namespace cpf {

// ---------------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------------
struct FftlogArgs {
  const double* in;
  double* out;
  const double* pre;       // [P, N]
  const double2* ut;       // fast: [P, N/2+1] u/N with Im dropped at DC/Nyquist (times (-1)^k for the pruned kernels);
                           // generic: [P, N] the same, Hermitian-extended to all N bins
  const double* post_re;   // [P, N]
  const double* post_im;   // [P, N] or null
  const double2* tw1;      // fast: [6, 256] factored twiddles (cpf_fft_core.h) ; generic: [N/2]
  const double2* tw2;      // fast: [6, 16]
  long long pairs_per_p;   // ceil(batch / 2)
  long long batch;
  int P, in_has_P, n, N, log2N, in_left, out_left, n_out, keep_padding;
  int ex_l_mode, ex_r_mode;
  double ex_l_val, ex_r_val;
  unsigned* tickets;       // ping-pong kernel with dynamic scheduling: one self-resetting ticket counter per plan row, else null
  unsigned* t_finished;    // ... the slot's CTA exit counter (self-resetting) and its completion word in mapped host memory
  unsigned* t_done;
  unsigned t_seq;
  int ahead;               // split kernel: a CTA prefetches into L2 the rows of the pair `ahead` pairs after its own (one wave of resident clusters)
};

// Last instruction of a dynamically scheduled kernel: the CTA that exits last publishes the launch's sequence number in mapped host
// memory, which is how the host knows that the slot's counters are back at zero before it hands the slot to another launch.
__device__ __forceinline__ void ticket_release(unsigned* finished, unsigned* done, const unsigned seq) {
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(finished, gridDim.x - 1) == gridDim.x - 1) *reinterpret_cast<volatile unsigned*>(done) = seq;
  }
}

// Two real rows ride one complex FFT, so a NaN / Inf in one row would spoil its partner, which the reference's row-wise
// numpy FFTs do not do.  Every kernel therefore replaces non-finite samples by zero on load, ORs "row a is bad" / "row b is
// bad" over the threads of the pair through two barriers it executes anyway (barrier.red), and writes NaN to the whole
// output row of a bad input row -- what numpy.fft returns for a row holding a NaN.
// (exponent field all ones, tested on the integer pipe: the fp64 pipe is the bottleneck of these kernels)
__device__ __forceinline__ bool nonfinite(const double v) { return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000; }
__device__ __forceinline__ double scrub(const double v, bool& bad) {
  if (nonfinite(v)) { bad = true; return 0.; }
  return v;
}
// The same for a thread's 8 + 8 samples of a pair in the persistent kernels: two integer instructions per sample on the
// common path (the minimum of ~hi & 0x7ff00000 is 0 iff some exponent field is all ones), the scrubbing itself in a branch
// that is almost never taken.
__device__ __forceinline__ void scrub_rows(double (&x)[8], double (&y)[8], bool& bad_a, bool& bad_b) {
  unsigned ma = 0x7ff00000u, mb = 0x7ff00000u;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    ma = min(ma, ~(unsigned)__double2hiint(x[r]) & 0x7ff00000u);
    mb = min(mb, ~(unsigned)__double2hiint(y[r]) & 0x7ff00000u);
  }
  bad_a = ma == 0u;
  bad_b = mb == 0u;
  if (bad_a | bad_b) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (nonfinite(x[r])) x[r] = 0.;
      if (nonfinite(y[r])) y[r] = 0.;
    }
  }
}

// value of the padded input at unpadded index i (i < 0 or i >= n is the extrapolated part) — fftlog.py:466-505
__device__ __forceinline__ double padded_value(const double* __restrict__ row, const int i, const FftlogArgs& a) {
  if ((unsigned)i < (unsigned)a.n) return __ldcs(row + i);
  if (i < 0) {
    if (a.ex_l_mode == CPF_EXTRAP_CONST) return a.ex_l_val;
    const double f0 = __ldg(row);
    if (a.ex_l_mode == CPF_EXTRAP_EDGE) return f0;
    return f0 * pow(__ldg(row + 1) / f0, (double)i);                    // :486-490
  }
  if (a.ex_r_mode == CPF_EXTRAP_CONST) return a.ex_r_val;
  const double fl = __ldg(row + a.n - 1);
  if (a.ex_r_mode == CPF_EXTRAP_EDGE) return fl;
  return fl / pow(__ldg(row + a.n - 2) / fl, (double)(i - (a.n - 1)));   // :497-501
}

__device__ __forceinline__ void store_out(const FftlogArgs& a, double* __restrict__ orow, const int o, double g,
                                          const double pr, const double pi, const bool cpost, const bool bad) {
  if (bad) g = nan("");
  if (cpost) {
    double2 r;
    r.x = g * pr;
    r.y = g * pi;
    __stcs(reinterpret_cast<double2*>(orow) + o, r);
  } else {
    __stcs(orow + o, g * pr);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fast path: N = 256*R1, one CTA of T = 16*R1 threads per pair of rows
// ---------------------------------------------------------------------------------------------------------------
template <int R1, bool PRUNED, bool CPOST>
__global__ void __launch_bounds__(16 * R1, 512 / (16 * R1)) fftlog_fast_kernel(const FftlogArgs a) {
  typedef Geo<R1> G;
  extern __shared__ double2 S[];
  constexpr int T = G::T, N = G::N;
  constexpr int NR = PRUNED ? 8 : 16;       // live registers rows on the pruned side
  constexpr int SHIFT = PRUNED ? N / 4 : 0;

  const int t = threadIdx.x;
  const long long q = blockIdx.x;
  const int p = (int)(q / a.pairs_per_p);            // p-major: CTAs resident together share one plan row's tables
  const long long b0 = 2 * (q - p * a.pairs_per_p), b1 = b0 + 1;
  const bool has1 = b1 < a.batch;
  const double* rowA = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
  const double* rowB = has1 ? a.in + (a.in_has_P ? (b1 * a.P + p) : b1) * (long long)a.n : rowA;
  const double* pre = a.pre + (size_t)p * N;

  double2 v[16];
  bool bad_a = false, bad_b = false;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int j = t + T * r + SHIFT;
    const int i = j - a.in_left;
    double x, y;
    if (PRUNED) {
      const bool ok = (unsigned)i < (unsigned)a.n;
      x = ok ? __ldcs(rowA + i) : 0.;
      y = (ok && has1) ? __ldcs(rowB + i) : 0.;
    } else {
      x = padded_value(rowA, i, a);
      y = has1 ? padded_value(rowB, i, a) : 0.;
    }
    x = scrub(x, bad_a);
    y = scrub(y, bad_b);
    const double pr = pre[j];
    v[r] = mk2(x * pr, y * pr);
  }
#pragma unroll
  for (int r = NR; r < 16; ++r) v[r] = mk2(0., 0.);

  // FFT #1
  fft_pass1<R1, PRUNED>(t, v, S, a.tw1);
  const bool row_a_bad = __syncthreads_or(bad_a);
  fft_pass2<R1>(t, S, a.tw2);
  const bool row_b_bad = __syncthreads_or(bad_b);
  fft_pass3<R1, false>(t, v, S);

  // kernel multiply: thread t holds bins k = t + T*r.  Only bins 0..N/2 are stored; k > N/2 uses conj(u[N-k]).
  // (plain loads, not __ldg: read-only-path loads may be hoisted above the barriers and then spill)
  const double2* uh = a.ut + (size_t)p * (N / 2 + 1);
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = cmul(v[r], uh[t + T * r]);
#pragma unroll
  for (int r = 8; r < 16; ++r) v[r] = cmul_conj(v[r], uh[T * (16 - r) - t]);
  __syncthreads();   // pass-3 reads of S are done before FFT #2 overwrites it

  // FFT #2
  fft_pass1<R1, false>(t, v, S, a.tw1);
  __syncthreads();
  fft_pass2<R1>(t, S, a.tw2);
  __syncthreads();
  fft_pass3<R1, PRUNED>(t, v, S);

  // un-bias, crop, store
  const size_t osz = (size_t)a.n_out * (CPOST ? 2 : 1);
  double* outA = a.out + (size_t)(b0 * a.P + p) * osz;
  double* outB = a.out + (size_t)(b1 * a.P + p) * osz;
  const double* post_re = a.post_re + (size_t)p * N;
  const double* post_im = CPOST ? a.post_im + (size_t)p * N : nullptr;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int j = t + T * r + SHIFT;
    const int o = a.keep_padding ? j : j - a.out_left;
    if ((unsigned)o < (unsigned)a.n_out) {
      const double pr = post_re[j];
      const double pi = CPOST ? post_im[j] : 0.;
      store_out(a, outA, o, v[r].x, pr, pi, CPOST, row_a_bad);
      if (has1) store_out(a, outB, o, v[r].y, pr, pi, CPOST, row_b_bad);
    }
  }
}

}  // namespace cpf

#include "cpf_fftlog_pp.cuh"
#include "cpf_fftlog_pp8k.cuh"
#include "cpf_fftlog_stream.cuh"

namespace cpf {

// ---------------------------------------------------------------------------------------------------------------
// N = 8192 (nk = 4096, cosmoprimo/fftlog.py:149-150): one CTA of 512 threads per pair of rows, two groups of 256 threads,
// each running the 4096-point register FFT on its own exchange buffer.  The 8192-point transforms are split by one
// radix-2 stage that needs no exchange:
//   FFT #1, decimation in frequency: A[2k + g] = FFT_4096( (a[n] + (-1)^g a[n + 4096]) w_8192^{g n} )[k]; group g owns the bins of parity g
//            (with zero padding a[n + 4096] = 0 after the N/4 rotation: both groups load the same 4096 samples);
//   FFT #2, decimation in time:      g[j] = E[j] + w_8192^j O[j],  g[j + 4096] = E[j] - w_8192^j O[j], E / O = FFT_4096 of the even / odd bins
//            (group 0 / group 1); the halves are exchanged through shared memory once, and with the cropped output only g[j], j < 4096,
//            is formed: group 0 finishes the elements of register rows 0..7, group 1 those of rows 8..15.
// ---------------------------------------------------------------------------------------------------------------
template <bool PRUNED, bool CPOST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2) fftlog_split2_kernel(const FftlogArgs a, const double2* __restrict__ tw8192) {
  typedef Geo<16> G;
  extern __shared__ double2 S[];
  constexpr int T = 256, M = 4096, N = 8192;
  constexpr int SHIFT = PRUNED ? N / 4 : 0;
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int g = (int)cluster.block_rank(), t = threadIdx.x;      // the two CTAs of a cluster are the two groups (two CTAs per SM: four groups resident)
  const double2* Sother = cluster.map_shared_rank(S, 1 - g);      // the other group's exchange buffer, through distributed shared memory
  const long long q = blockIdx.x >> 1;
  const int p = (int)(q / a.pairs_per_p);
  const long long b0 = 2 * (q - p * a.pairs_per_p), b1 = b0 + 1;
  const bool has1 = b1 < a.batch;
  const double* rowA = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
  const double* rowB = has1 ? a.in + (a.in_has_P ? (b1 * a.P + p) : b1) * (long long)a.n : rowA;
  const double* pre = a.pre + (size_t)p * N;

  // L2 prefetch of the two rows of a pair that a later wave of clusters transforms (group 0: row a, group 1: row b; one 128-byte line per
  // thread covers a row of up to 4096 samples): the input loads of that wave then hit L2 instead of HBM
  if (a.ahead > 0) {
    const long long qn = q + a.ahead;
    if (qn / a.pairs_per_p == p) {                       // same plan row: same input indexing
      const long long bn = 2 * (qn - p * a.pairs_per_p) + g;
      if (bn < a.batch && 16 * t < a.n) {
        const double* rn = a.in + (a.in_has_P ? (bn * a.P + p) : bn) * (long long)a.n + 16 * t;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rn));
      }
    }
  }
  double2 v[16];
  bool bad_a = false, bad_b = false;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int nn = t + T * r;                          // n < 4096
    const int j = nn + SHIFT, i = j - a.in_left;
    double x, y;
    if (PRUNED) {
      const bool ok = (unsigned)i < (unsigned)a.n;
      x = ok ? __ldcs(rowA + i) : 0.;
      y = (ok && has1) ? __ldcs(rowB + i) : 0.;
    } else {
      x = padded_value(rowA, i, a);
      y = has1 ? padded_value(rowB, i, a) : 0.;
    }
    x = scrub(x, bad_a);
    y = scrub(y, bad_b);
    const double pr = pre[j];
    double2 z = mk2(x * pr, y * pr);
    if (!PRUNED) {                                     // upper half of the padded row
      const int j2 = j + M, i2 = j2 - a.in_left;
      const double x2 = scrub(padded_value(rowA, i2, a), bad_a), y2 = has1 ? scrub(padded_value(rowB, i2, a), bad_b) : 0.;
      const double pr2 = pre[j2];
      z = g == 0 ? mk2(fma(x2, pr2, z.x), fma(y2, pr2, z.y)) : mk2(fma(-x2, pr2, z.x), fma(-y2, pr2, z.y));
    }
    v[r] = g == 0 ? z : cmul(z, __ldg(tw8192 + nn));
  }
  // FFT #1: the groups run free of each other until the final exchange
  fft_pass1<16, false>(t, v, S, a.tw1);
  const bool row_a_bad = __syncthreads_or(bad_a);                 // both groups loaded the same rows: the same flags in both
  fft_pass2<16>(t, S, a.tw2);
  const bool row_b_bad = __syncthreads_or(bad_b);
  fft_pass3<16, false>(t, v, S);
  // kernel multiply: this thread holds bins 2 (t + 256 r) + g; bins above N/2 use conj(u[N - bin])
  const double2* uh = a.ut + (size_t)p * (N / 2 + 1);
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = cmul(v[r], uh[2 * (t + T * r) + g]);
#pragma unroll
  for (int r = 8; r < 16; ++r) v[r] = cmul_conj(v[r], uh[N - 2 * (t + T * r) - g]);
  __syncthreads();
  // FFT #2
  fft_pass1<16, false>(t, v, S, a.tw1);
  __syncthreads();
  fft_pass2<16>(t, S, a.tw2);
  __syncthreads();
  fft_pass3<16, false>(t, v, S);
  if (g == 1) {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = cmul(v[r], __ldg(tw8192 + t + T * r));       // O'[j] = w_8192^j O[j]
  }
  __syncthreads();                                      // this group's pass-3 reads of S are done
  // hand the other group what it needs: cropped output: group 0 finishes rows r < 8, group 1 rows r >= 8; otherwise both need everything
#pragma unroll
  for (int r = 0; r < 16; ++r)
    if (!PRUNED || (g == 0 ? r >= 8 : r < 8)) S[t + T * r] = v[r];
  cluster.barrier_arrive();                             // this group's half is written ...
  const size_t osz = (size_t)a.n_out * (CPOST ? 2 : 1);
  double* outA = a.out + (size_t)(b0 * a.P + p) * osz;
  double* outB = a.out + (size_t)(b1 * a.P + p) * osz;
  const double* post_re = a.post_re + (size_t)p * N;
  const double* post_im = CPOST ? a.post_im + (size_t)p * N : nullptr;
  cluster.barrier_wait();                               // ... and so is the other group's, visible across the cluster
  double2 wo[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (PRUNED && (g == 0 ? r >= 8 : r < 8)) continue;
    wo[r] = Sother[t + T * r];
  }
  cluster.barrier_arrive();                             // done reading the other group's shared memory (waited for at the very end)
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (PRUNED && (g == 0 ? r >= 8 : r < 8)) continue;
    const double2 w = wo[r];
    double2 res;
    int j;
    if (PRUNED) { res = mk2(v[r].x + w.x, v[r].y + w.y); j = t + T * r + SHIFT; }        // E + O' either way
    else if (g == 0) { res = mk2(v[r].x + w.x, v[r].y + w.y); j = t + T * r; }            // g[j] = E + O'
    else { res = mk2(w.x - v[r].x, w.y - v[r].y); j = t + T * r + M; }                    // g[j + 4096] = E - O'
    const int o = a.keep_padding ? j : j - a.out_left;
    if ((unsigned)o < (unsigned)a.n_out) {
      const double pr = post_re[j];
      const double pi = CPOST ? post_im[j] : 0.;
      store_out(a, outA, o, res.x, pr, pi, CPOST, row_a_bad);
      if (has1) store_out(a, outB, o, res.y, pr, pi, CPOST, row_b_bad);
    }
  }
  cluster.barrier_wait();                               // neither CTA leaves while the other still reads its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// generic path: any power-of-two N <= CPF_MAX_N, whole transform in shared memory, radix-2
// ---------------------------------------------------------------------------------------------------------------
// forward FFT, natural order in and out; tw[k] = exp(-2 pi i k/N), k < N/2
__device__ void block_fft_forward(double2* S, const int N, const int log2N, const double2* __restrict__ tw) {
  const int half = N >> 1;
  for (int h = half; h >= 1; h >>= 1) {       // decimation in frequency
    const int step = half / h;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const int lo = i & (h - 1);
      const int j = ((i - lo) << 1) | lo;
      const double2 x0 = S[j], x1 = S[j + h];
      S[j] = mk2(x0.x + x1.x, x0.y + x1.y);
      S[j + h] = cmul(mk2(x0.x - x1.x, x0.y - x1.y), __ldg(tw + lo * step));
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) {   // undo the bit reversal
    const int r = (int)(__brev((unsigned)i) >> (32 - log2N));
    if (i < r) {
      const double2 tmp = S[i];
      S[i] = S[r];
      S[r] = tmp;
    }
  }
  __syncthreads();
}

template <bool CPOST>
__global__ void fftlog_generic_kernel(const FftlogArgs a) {
  extern __shared__ double2 S[];
  const int N = a.N;
  const long long q = blockIdx.x;
  const int p = (int)(q / a.pairs_per_p);            // p-major: CTAs resident together share one plan row's tables
  const long long b0 = 2 * (q - p * a.pairs_per_p), b1 = b0 + 1;
  const bool has1 = b1 < a.batch;
  const double* rowA = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
  const double* rowB = has1 ? a.in + (a.in_has_P ? (b1 * a.P + p) : b1) * (long long)a.n : rowA;
  const double* pre = a.pre + (size_t)p * N;
  bool bad_a = false, bad_b = false;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const int i = j - a.in_left;
    const double pr = pre[j];
    const double x = scrub(padded_value(rowA, i, a), bad_a), y = has1 ? scrub(padded_value(rowB, i, a), bad_b) : 0.;
    S[j] = mk2(x * pr, y * pr);
  }
  const bool row_a_bad = __syncthreads_or(bad_a);
  block_fft_forward(S, N, a.log2N, a.tw1);
  const double2* ut = a.ut + (size_t)p * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) S[j] = cmul(S[j], __ldg(ut + j));
  const bool row_b_bad = __syncthreads_or(bad_b);
  block_fft_forward(S, N, a.log2N, a.tw1);
  const size_t osz = (size_t)a.n_out * (CPOST ? 2 : 1);
  double* outA = a.out + (size_t)(b0 * a.P + p) * osz;
  double* outB = a.out + (size_t)(b1 * a.P + p) * osz;
  const double* post_re = a.post_re + (size_t)p * N;
  const double* post_im = CPOST ? a.post_im + (size_t)p * N : nullptr;
  const int off = a.keep_padding ? 0 : a.out_left;
  for (int o = threadIdx.x; o < a.n_out; o += blockDim.x) {
    const int j = o + off;
    const double pr = __ldg(post_re + j);
    const double pi = CPOST ? __ldg(post_im + j) : 0.;
    store_out(a, outA, o, S[j].x, pr, pi, CPOST, row_a_bad);
    if (has1) store_out(a, outB, o, S[j].y, pr, pi, CPOST, row_b_bad);
  }
}

// ---- unfused engine duck type (NumpyFFTEngine.forward/backward, fftlog.py:538-544) ----------------------------
// rfft of two packed rows: Z = FFT(a + i b); A[m] = (Z[m] + conj Z[N-m])/2, B[m] = (Z[m] - conj Z[N-m])/(2i)
__global__ void rfft_pair_kernel(const double* __restrict__ in, double2* __restrict__ out, const long long rows,
                                 const int N, const int log2N, const double2* __restrict__ tw) {
  extern __shared__ double2 S[];
  const long long b0 = 2LL * blockIdx.x, b1 = b0 + 1;
  const bool has1 = b1 < rows;
  const double* rowA = in + b0 * N;
  const double* rowB = in + (has1 ? b1 : b0) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) S[j] = mk2(rowA[j], has1 ? rowB[j] : 0.);
  __syncthreads();
  block_fft_forward(S, N, log2N, tw);
  const int nb = N / 2 + 1;
  for (int m = threadIdx.x; m < nb; m += blockDim.x) {
    const double2 z = S[m], zc = S[(N - m) & (N - 1)];
    out[b0 * nb + m] = mk2(0.5 * (z.x + zc.x), 0.5 * (z.y - zc.y));
    if (has1) out[b1 * nb + m] = mk2(0.5 * (z.y + zc.y), 0.5 * (zc.x - z.x));
  }
}

// irfft(conj(X), n=N) of two packed rows: W = Xa~ + i Xb~ (Hermitian extensions, Im dropped at DC/Nyquist);
// FFT_fwd(W)/N = ga + i gb
__global__ void irfft_conj_pair_kernel(const double2* __restrict__ in, double* __restrict__ out, const long long rows,
                                       const int N, const int log2N, const double2* __restrict__ tw) {
  extern __shared__ double2 S[];
  const long long b0 = 2LL * blockIdx.x, b1 = b0 + 1;
  const bool has1 = b1 < rows;
  const int nb = N / 2 + 1;
  const double2* rowA = in + b0 * nb;
  const double2* rowB = in + (has1 ? b1 : b0) * nb;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const int m = k <= N / 2 ? k : N - k;
    double2 xa = rowA[m], xb = has1 ? rowB[m] : mk2(0., 0.);
    if (k > N / 2) { xa.y = -xa.y; xb.y = -xb.y; }
    if (m == 0 || m == N / 2) { xa.y = 0.; xb.y = 0.; }
    S[k] = mk2(xa.x - xb.y, xa.y + xb.x);
  }
  __syncthreads();
  block_fft_forward(S, N, log2N, tw);
  const double inv = 1. / N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    out[b0 * N + j] = S[j].x * inv;
    if (has1) out[b1 * N + j] = S[j].y * inv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static double2 unit_root(long long num, long long den) {   // exp(-2 pi i num/den), long double evaluation
  num %= den;
  const long double ang = -2.0L * acosl(-1.0L) * (long double)num / (long double)den;
  double2 r;
  r.x = (double)cosl(ang);
  r.y = (double)sinl(ang);
  // exact values on the axes
  if ((4 * num) % den == 0) {
    const long long qd = (4 * num) / den;
    r.x = qd == 0 ? 1. : (qd == 2 ? -1. : 0.);
    r.y = qd == 1 ? -1. : (qd == 3 ? 1. : 0.);
  }
  return r;
}

std::atomic<long long> g_stat_dynamic{0}, g_stat_fallback{0};      // cpf_counter

// private scratch pool per device (see cpf_common.h)
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {};

cudaMemPool_t scratch_pool(int device) {
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (g_pools[device]) return g_pools[device];
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  cudaMemPool_t pool = nullptr;
  if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  uint64_t keep = 2048ull << 20;
  if (const char* e = getenv("CPF_SCRATCH_KEEP_MB")) { const long long v = atoll(e); if (v >= 0) keep = (uint64_t)v << 20; }
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  g_pools[device] = pool;
  return pool;
}

// One-time table upload.  cudaMemcpy from pageable memory returns once the data sits in the driver's staging buffer: the DMA into *dptr
// may still be in flight on the legacy stream, and the kernels that read the table run on non-blocking streams (the staging pool's, the
// caller's), which do not wait for it.  So the legacy stream is drained before the table is handed out (r4o: the first chunks of a
// host-buffer call read a table that had not landed yet, once the host side stopped being slow enough to hide it).
int upload(void** dptr, const void* src, size_t bytes) {
  CPF_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
  CPF_CUDA(cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
  CPF_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  return CPF_OK;
}

static int fast_radix(int N) { return N == 4096 ? 16 : N == 2048 ? 8 : N == 1024 ? 4 : 0; }

// per-thread table records of the ping-pong kernel (layout: cpf_fftlog_pp.cuh).  uhs = (-1)^k u/N, [P, N/2+1];
// tw = the batch-invariant part of a record ([T][32], see FastTw)
static void build_pp_tables(int N, int R1, int P, const double* pre, const std::vector<double2>& uhs, const double* post_re,
                            const std::vector<double2>& tw, std::vector<double2>& tab) {
  const int T = 16 * R1, nb = N / 2 + 1;
  double2 zero; zero.x = 0.; zero.y = 0.;
  tab.assign((size_t)P * T * PP_REC, zero);
  for (int p = 0; p < P; ++p)
    for (int t = 0; t < T; ++t) {
      double2* rec = &tab[((size_t)p * T + t) * PP_REC];
      for (int e = 0; e < 32; ++e) rec[e] = tw[(size_t)t * 32 + e];
      for (int r = 0; r < 16; ++r) {
        const int k = t + T * r;
        double2 v = uhs[(size_t)p * nb + (k <= N / 2 ? k : N - k)];
        if (k > N / 2) v.y = -v.y;
        rec[32 + r] = v;
      }
      double* dpre = reinterpret_cast<double*>(rec + 48);
      double* dpost = reinterpret_cast<double*>(rec + 52);
      for (int r = 0; r < 8; ++r) {
        const size_t j = (size_t)p * N + N / 4 + t + T * r;
        dpre[r] = pre[j];
        dpost[r] = post_re[j];
      }
    }
}

// records of fftlog_pp8k_kernel (cpf_fftlog_pp8k.cuh): N = 8192 as two 4096-point chains; pp_tw are the ping-pong twiddles of N = 4096
static void build_pp8k_tables(int P, const double* pre, const std::vector<double2>& uhs, const double* post_re, const std::vector<double2>& pp_tw,
                              std::vector<double2>& tab) {
  const int N = 8192, T = 256, nb = N / 2 + 1;
  double2 zero; zero.x = 0.; zero.y = 0.;
  tab.assign((size_t)P * T * P8_REC, zero);
  for (int p = 0; p < P; ++p)
    for (int t = 0; t < T; ++t) {
      double2* rec = &tab[((size_t)p * T + t) * P8_REC];
      for (int k1 = 0; k1 < 16; ++k1) rec[k1] = pp_tw[(size_t)t * 32 + k1];
      for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 16; ++r) {
          const int b = 2 * (t + T * r) + c;
          double2 v = uhs[(size_t)p * nb + (b <= N / 2 ? b : N - b)];
          if (b > N / 2) v.y = -v.y;
          rec[16 + 16 * c + r] = v;
        }
      double* dpre = reinterpret_cast<double*>(rec + 48);
      double* dpost = reinterpret_cast<double*>(rec + 56);
      for (int r = 0; r < 16; ++r) {
        const size_t j = (size_t)p * N + N / 4 + t + T * r;
        dpre[r] = pre[j];
        dpost[r] = post_re[j];
      }
    }
}

// kernel spectrum in the order fftlog_stream_kernel reads it (cpf_fftlog_stream.cuh): [P][16][256], thread tau = 16 H + L
// holds the bins H + 16 L + 256 l2
static void build_stream_ut(int P, const std::vector<double2>& uhs, double2* ut) {
  const int N = 4096, T = 256, nb = N / 2 + 1;
  for (int p = 0; p < P; ++p)
    for (int t = 0; t < T; ++t)
      for (int k = 0; k < 16; ++k) {
        const int bin = (t >> 4) + 16 * (t & 15) + 256 * k;
        double2 v = uhs[(size_t)p * nb + (bin <= N / 2 ? bin : N - bin)];
        if (bin > N / 2) v.y = -v.y;
        ut[((size_t)p * 16 + k) * T + t] = v;
      }
}

// Batch- and plan-invariant twiddle tables of the register kernels, built once per (device, N) and never freed.
struct FastTw {
  int device, N;
  double2* d_tw1;                 // per-pair kernel: [6, 256] factored pass-1 twiddles (cpf_fft_core.h)
  double2* d_tw2;                 // per-pair kernel: [6, 16]
  double* d_st_tw;                // stream kernel (N = 4096): scaled twiddles + ratios of P1 / P2 / P1', [3, 32, 256] (st_build_tables)
  double* d_m256;                 // stream kernel: P2' twiddles and P3' ratios, [2, 16, 16]
  std::vector<double2> pp_tw;     // host: twiddle part of the ping-pong records, [T][32]
};
static std::mutex g_fast_mutex;
static std::vector<FastTw*> g_fast_cache;

static int fast_twiddles(int device, int N, const FastTw** out) {
  std::lock_guard<std::mutex> lock(g_fast_mutex);
  for (auto* e : g_fast_cache)
    if (e->device == device && e->N == N) { *out = e; return CPF_OK; }
  const int R1 = N / 256, T = 16 * R1, C = 16 / R1;
  FastTw* f = new FastTw();
  f->device = device; f->N = N;
  f->d_tw1 = f->d_tw2 = nullptr;
  f->d_st_tw = f->d_m256 = nullptr;
  static const int expo[6] = {1, 2, 3, 4, 8, 12};
  std::vector<double2> tw1(6 * 256), tw2(6 * 16);
  for (int e = 0; e < 6; ++e) {
    for (int n2 = 0; n2 < 256; ++n2) tw1[e * 256 + n2] = unit_root((long long)expo[e] * n2, N);
    for (int m2 = 0; m2 < 16; ++m2) tw2[e * 16 + m2] = unit_root(expo[e] * m2, 256);
  }
  f->pp_tw.resize((size_t)T * 32);
  for (int t = 0; t < T; ++t) {
    for (int c = 0; c < C; ++c)
      for (int k1 = 0; k1 < R1; ++k1) f->pp_tw[(size_t)t * 32 + c * R1 + k1] = unit_root((long long)(t + T * c) * k1, N);
    for (int l1 = 0; l1 < 16; ++l1) f->pp_tw[(size_t)t * 32 + 16 + l1] = unit_root((long long)(t & 15) * l1, 256);
  }
  int rc = upload((void**)&f->d_tw1, tw1.data(), tw1.size() * sizeof(double2));
  if (rc == CPF_OK) rc = upload((void**)&f->d_tw2, tw2.data(), tw2.size() * sizeof(double2));
  if (rc == CPF_OK && N == 4096) {
    // scaled twiddles + butterfly ratios of the stream kernel (cpf_stream_core.h)
    std::vector<double> stw((size_t)3 * 32 * 256), m256(512);
    st_build_tables(stw.data(), m256.data());
    rc = upload((void**)&f->d_st_tw, stw.data(), stw.size() * sizeof(double));
    if (rc == CPF_OK) rc = upload((void**)&f->d_m256, m256.data(), m256.size() * sizeof(double));
  }
  if (rc != CPF_OK) {
    cudaFree(f->d_tw1); cudaFree(f->d_tw2); cudaFree(f->d_st_tw); cudaFree(f->d_m256);
    delete f;
    return rc;
  }
  g_fast_cache.push_back(f);
  *out = f;
  return CPF_OK;
}

// w_8192^n, n < 4096, of the N = 8192 kernels and the pass-2 twiddles w_256^{m2 l1} ([l1][m2]) of fftlog_pp8k_kernel: once per device, never freed
struct Tw8192 { int device; double2 *tw, *tw2; };
static std::mutex g_tw8192_mutex;
static std::vector<Tw8192> g_tw8192;
static int split2_twiddles(int device, const double2** out, const double2** out_tw2) {
  std::lock_guard<std::mutex> lock(g_tw8192_mutex);
  for (auto& e : g_tw8192)
    if (e.device == device) { *out = e.tw; *out_tw2 = e.tw2; return CPF_OK; }
  std::vector<double2> tw(4096), tw2(256);
  for (int n = 0; n < 4096; ++n) tw[n] = unit_root(n, 8192);
  for (int l1 = 0; l1 < 16; ++l1)
    for (int m2 = 0; m2 < 16; ++m2) tw2[16 * l1 + m2] = unit_root(m2 * l1, 256);
  Tw8192 e;
  e.device = device; e.tw = e.tw2 = nullptr;
  CPF_TRY(upload((void**)&e.tw, tw.data(), tw.size() * sizeof(double2)));
  CPF_TRY(upload((void**)&e.tw2, tw2.data(), tw2.size() * sizeof(double2)));
  g_tw8192.push_back(e);
  *out = e.tw; *out_tw2 = e.tw2;
  return CPF_OK;
}

// per-device twiddle cache for the unfused engine entry points
struct GenericTw {
  int device, N;
  double2* d_tw;
};
static std::mutex g_tw_mutex;
static std::vector<GenericTw> g_tw_cache;

static int generic_twiddles(int device, int N, double2** out) {
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  for (auto& e : g_tw_cache)
    if (e.device == device && e.N == N) { *out = e.d_tw; return CPF_OK; }
  std::vector<double2> tw(N / 2 > 0 ? N / 2 : 1);
  for (int k = 0; k < N / 2; ++k) tw[k] = unit_root(k, N);
  if (N < 2) tw[0] = unit_root(0, 1);
  double2* d = nullptr;
  CPF_TRY(upload((void**)&d, tw.data(), tw.size() * sizeof(double2)));
  g_tw_cache.push_back({device, N, d});
  *out = d;
  return CPF_OK;
}

// staging streams for host-pointer calls (H2D / kernel / D2H of successive chunks overlap)
static const int kNumStage = 8;
struct StagePool {
  int device;
  cudaStream_t h2d, comp, d2h;      // one stream per engine: the two copy queues never wait behind a kernel
  cudaEvent_t in_ready[kNumStage];  // H2D of the chunk in buffer i has landed
  cudaEvent_t k_done[kNumStage];    // kernel on buffer i finished (input buffer i is free, output buffer i is full)
  cudaEvent_t out_free[kNumStage];  // D2H of buffer i finished
  cudaEvent_t start;
  std::mutex busy;                  // host-pointer calls on one device are serialised
  void* h_bounce[kNumStage] = {};   // page-locked bounce buffers for pageable input (HostCopyPool), grown on demand
  size_t h_bounce_bytes = 0;
  void* h_small[2] = {};            // mapped page-locked buffers of the small-call path (kSmallCallBytes each): in, out
  void* d_small[2] = {};            // ... and their device addresses
};
constexpr size_t kSmallCallBytes = 256u << 10;
static std::mutex g_stage_mutex;
static std::vector<StagePool*> g_stage_pools;

static int stage_pool(int device, StagePool** out) {
  std::lock_guard<std::mutex> lock(g_stage_mutex);
  for (auto* e : g_stage_pools)
    if (e->device == device) { *out = e; return CPF_OK; }
  StagePool* sp = new StagePool();
  sp->device = device;
  CPF_CUDA(cudaStreamCreateWithFlags(&sp->h2d, cudaStreamNonBlocking));
  CPF_CUDA(cudaStreamCreateWithFlags(&sp->comp, cudaStreamNonBlocking));
  CPF_CUDA(cudaStreamCreateWithFlags(&sp->d2h, cudaStreamNonBlocking));
  for (int i = 0; i < kNumStage; ++i) {
    CPF_CUDA(cudaEventCreateWithFlags(&sp->in_ready[i], cudaEventDisableTiming));
    CPF_CUDA(cudaEventCreateWithFlags(&sp->k_done[i], cudaEventDisableTiming));
    CPF_CUDA(cudaEventCreateWithFlags(&sp->out_free[i], cudaEventDisableTiming));
  }
  CPF_CUDA(cudaEventCreateWithFlags(&sp->start, cudaEventDisableTiming));
  g_stage_pools.push_back(sp);
  *out = sp;
  return CPF_OK;
}

}  // namespace cpf

using namespace cpf;

struct cpf_plan {
  int n, N, P, in_left, out_left, device, log2N;
  bool post_complex;
  int fast_R1;
  bool split2 = false;        // N = 8192: two 4096-point register FFTs per transform (fftlog_split2_kernel)
  const double2* d_tw8192 = nullptr;   // w_8192^n, n < 4096 (shared, not owned)
  const double2* d_tw2_256 = nullptr;  // w_256^{m2 l1}, [16][16] (shared, not owned)
  mutable double2* d_pp8k = nullptr;   // N = 8192: per-thread records of fftlog_pp8k_kernel [P, 256, P8_REC], built on first use
  bool window_prunable;
  void* d_block = nullptr;    // one device allocation holds every per-plan table below
  double* d_pre = nullptr;
  double2* d_ut = nullptr;    // generic plans: [P,N] Hermitian-extended; fast plans: [P,N/2+1]
  double2* d_uts = nullptr;   // fast plans: (-1)^k u/N  (input and output windows rotated by N/4)
  double* d_post_re = nullptr;
  double* d_post_im = nullptr;
  double2* d_tw = nullptr;    // generic [N/2]
  double2* d_st_ut = nullptr; // stream kernel: kernel spectrum per thread [P, 16, 256]
  const FastTw* fast = nullptr;   // batch-invariant twiddles of the register kernels (shared, not owned)
  // ping-pong kernel: per-thread table records [P, T, PP_REC] (cpf_fftlog_pp.cuh); built at plan creation for
  // N = 2048 / 1024, on first use for N = 4096 (where the stream kernel is the default)
  mutable double2* d_pp = nullptr;
  mutable std::mutex pp_mutex;
  std::vector<double> h_pre_win, h_post_win;   // host copies kept for the lazy ping-pong tables (N = 4096 only)
  std::vector<double2> h_uhs;
};

extern "C" {

int cpf_version(void) { return CPF_VERSION; }

const char* cpf_last_error(void) { return g_last_error.c_str(); }

int cpf_device_count(int* count) {
  if (!count) return fail(CPF_EINVAL, "cpf_device_count: null pointer");
  *count = 0;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) return fail(CPF_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  *count = c;
  return CPF_OK;
}

int64_t cpf_counter(int which) {
  switch (which) {
    case CPF_COUNTER_DYNAMIC_LAUNCHES: return cpf::g_stat_dynamic.load();
    case CPF_COUNTER_TICKET_FALLBACKS: return cpf::g_stat_fallback.load();
    default: return -1;
  }
}

int cpf_trim(int device) {
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (device >= 0 && device < 64) pool = g_pools[device];
  }
  if (!pool) return CPF_OK;             // nothing was ever allocated on that device
  DeviceGuard guard(device);
  CPF_CUDA(cudaDeviceSynchronize());
  CPF_CUDA(cudaMemPoolTrimTo(pool, 0));
  return CPF_OK;
}

int cpf_plan_destroy(cpf_plan* plan) {
  if (!plan) return CPF_OK;
  DeviceGuard guard(plan->device);
  cudaFree(plan->d_block);
  cudaFree(plan->d_pp);
  cudaFree(plan->d_pp8k);
  delete plan;
  return CPF_OK;
}

int cpf_plan_create(cpf_plan** out, int n, int N, int P, int in_left, int out_left, const double* pre,
                    const double* u_ri, const double* post_re, const double* post_im, int device) {
  if (!out) return fail(CPF_EINVAL, "cpf_plan_create: null plan pointer");
  *out = nullptr;
  if (!pre || !u_ri || !post_re) return fail(CPF_EINVAL, "cpf_plan_create: null table pointer");
  if (n < 1 || P < 1) return fail(CPF_EINVAL, "cpf_plan_create: n=%d, P=%d must be positive", n, P);
  if (!is_pow2(N) || N < 2) return fail(CPF_EINVAL, "cpf_plan_create: padded size N=%d must be a power of two >= 2", N);
  if (N > CPF_MAX_N) return fail(CPF_EUNSUPPORTED, "cpf_plan_create: padded size N=%d exceeds CPF_MAX_N=%d", N, CPF_MAX_N);
  if (n > N || in_left < 0 || out_left < 0 || in_left + n > N || out_left + n > N)
    return fail(CPF_EINVAL, "cpf_plan_create: inconsistent sizes n=%d N=%d in_left=%d out_left=%d", n, N, in_left, out_left);
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_plan_create: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return fail(CPF_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(guard.err));

  cpf_plan* pl = new cpf_plan();
  pl->n = n; pl->N = N; pl->P = P; pl->in_left = in_left; pl->out_left = out_left; pl->device = device;
  pl->log2N = ilog2(N);
  pl->post_complex = post_im != nullptr;
  pl->fast_R1 = fast_radix(N);
  pl->split2 = (N == 8192);
  pl->window_prunable = in_left >= N / 4 && in_left + n <= 3 * (N / 4) && out_left >= N / 4 && out_left + n <= 3 * (N / 4);

  const size_t PN = (size_t)P * N;
  const int nb = N / 2 + 1;
  const double inv = 1. / N;
  // u/N with the imaginary parts that numpy.fft.irfft discards (DC, Nyquist) removed
  std::vector<double2> uh((size_t)P * nb);
  for (int p = 0; p < P; ++p)
    for (int m = 0; m < nb; ++m) {
      double2 v;
      v.x = u_ri[2 * ((size_t)p * nb + m)] * inv;
      v.y = (m == 0 || m == N / 2) ? 0. : u_ri[2 * ((size_t)p * nb + m) + 1] * inv;
      uh[(size_t)p * nb + m] = v;
    }
  // every per-plan table goes into one host staging buffer and one device allocation (one cudaMalloc + one copy)
  std::vector<char> stage;
  auto put = [&stage](const void* src, size_t bytes) {
    const size_t off = (stage.size() + 255) & ~(size_t)255;
    stage.resize(off + bytes);
    if (src) memcpy(stage.data() + off, src, bytes);
    return off;
  };
  const size_t NONE = ~(size_t)0;
  size_t o_pre = put(pre, PN * sizeof(double)), o_post = put(post_re, PN * sizeof(double));
  size_t o_post_im = post_im ? put(post_im, PN * sizeof(double)) : NONE;
  size_t o_ut = NONE, o_uts = NONE, o_st_ut = NONE;
  std::vector<double2> pp_tab;
  int rc = CPF_OK;
  do {
    if (pl->fast_R1 || pl->split2) {
      if ((rc = fast_twiddles(device, pl->split2 ? 4096 : N, &pl->fast))) break;
      if (pl->split2 && (rc = split2_twiddles(device, &pl->d_tw8192, &pl->d_tw2_256))) break;
      std::vector<double2> uhs(uh);
      for (int p = 0; p < P; ++p)
        for (int m = 1; m < nb; m += 2) { uhs[(size_t)p * nb + m].x = -uhs[(size_t)p * nb + m].x; uhs[(size_t)p * nb + m].y = -uhs[(size_t)p * nb + m].y; }
      o_ut = put(uh.data(), uh.size() * sizeof(double2));
      o_uts = put(uhs.data(), uhs.size() * sizeof(double2));
      if (pl->fast_R1 && pl->window_prunable && !post_im) {
        if (pl->fast_R1 == 16) {
          o_st_ut = put(nullptr, (size_t)P * 16 * 256 * sizeof(double2));
          build_stream_ut(P, uhs, reinterpret_cast<double2*>(stage.data() + o_st_ut));
          pl->h_uhs.swap(uhs);                      // kept for the lazy ping-pong tables
          pl->h_pre_win.assign(pre, pre + PN);
          pl->h_post_win.assign(post_re, post_re + PN);
        } else {
          build_pp_tables(N, pl->fast_R1, P, pre, uhs, post_re, pl->fast->pp_tw, pp_tab);
        }
      } else if (pl->split2 && pl->window_prunable && !post_im) {
        pl->h_uhs.swap(uhs);                        // kept for the lazy records of the persistent N = 8192 kernel
        pl->h_pre_win.assign(pre, pre + PN);
        pl->h_post_win.assign(post_re, post_re + PN);
      }
    } else {
      std::vector<double2> ut(PN);
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < N; ++k) {
          double2 v = uh[(size_t)p * nb + (k <= N / 2 ? k : N - k)];
          if (k > N / 2) v.y = -v.y;
          ut[(size_t)p * N + k] = v;
        }
      o_ut = put(ut.data(), ut.size() * sizeof(double2));
      double2* tw = nullptr;
      if ((rc = generic_twiddles(device, N, &tw))) break;
      pl->d_tw = tw;
    }
    if ((rc = upload(&pl->d_block, stage.data(), stage.size()))) break;
    if (!pp_tab.empty() && (rc = upload((void**)&pl->d_pp, pp_tab.data(), pp_tab.size() * sizeof(double2)))) break;
    char* base = static_cast<char*>(pl->d_block);
    pl->d_pre = reinterpret_cast<double*>(base + o_pre);
    pl->d_post_re = reinterpret_cast<double*>(base + o_post);
    if (o_post_im != NONE) pl->d_post_im = reinterpret_cast<double*>(base + o_post_im);
    if (o_ut != NONE) pl->d_ut = reinterpret_cast<double2*>(base + o_ut);
    if (o_uts != NONE) pl->d_uts = reinterpret_cast<double2*>(base + o_uts);
    if (o_st_ut != NONE) pl->d_st_ut = reinterpret_cast<double2*>(base + o_st_ut);
  } while (0);
  if (rc != CPF_OK) {
    std::string keep = g_last_error;
    cpf_plan_destroy(pl);
    g_last_error = keep;
    return rc;
  }
  *out = pl;
  return CPF_OK;
}

static bool use_pruned(const cpf_plan* pl, int ex_l_mode, double ex_l_val, int ex_r_mode, double ex_r_val, int keep_padding) {
  return (pl->fast_R1 || pl->split2) && pl->window_prunable && !keep_padding && ex_l_mode == CPF_EXTRAP_CONST && ex_r_mode == CPF_EXTRAP_CONST &&
         ex_l_val == 0. && ex_r_val == 0.;
}

int cpf_plan_kernel_family(const cpf_plan* plan, int ex_l_mode, double ex_l_val, int ex_r_mode, double ex_r_val, int keep_padding) {
  if (!plan) return -1;
  if (!plan->fast_R1 && !plan->split2) return 0;
  return use_pruned(plan, ex_l_mode, ex_l_val, ex_r_mode, ex_r_val, keep_padding) ? 2 : 1;
}

}  // extern "C"

namespace cpf {

template <int R1, bool PRUNED, bool CPOST>
static int launch_fast(const FftlogArgs& a, long long nblocks, cudaStream_t stream) {
  typedef Geo<R1> G;
  const size_t smem = (size_t)G::SMEM_ELEMS * sizeof(double2);
  auto kern = fftlog_fast_kernel<R1, PRUNED, CPOST>;
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)nblocks, G::T, smem, stream>>>(a);
  CPF_CUDA(cudaGetLastError());
  return CPF_OK;
}

template <int R1>
static int launch_fast_r(const FftlogArgs& a, bool pruned, bool cpost, long long nblocks, cudaStream_t stream) {
  if (pruned) return cpost ? launch_fast<R1, true, true>(a, nblocks, stream) : launch_fast<R1, true, false>(a, nblocks, stream);
  return cpost ? launch_fast<R1, false, true>(a, nblocks, stream) : launch_fast<R1, false, false>(a, nblocks, stream);
}

// Ticket counters of the dynamically scheduled persistent kernels: a per-device ring of slots, each with kTicketSlotWords counters (one
// per plan row), a CTA exit counter and a completion word in mapped host memory.  A launch takes the next slot of the ring; its counters
// are back at zero when it completes (st_draw_ticket).  Before a slot is handed out again the host checks that its previous launch has
// published its sequence number (ticket_release): a slot whose launch is still queued or running -- more launches in flight than slots,
// whatever the streams -- or never finished is NOT reused, and the new launch takes the static split instead.  No reset kernel, no
// synchronisation, one host read per launch.  CPF_TICKET_SLOTS (1..64, read once) shrinks the ring (tests).
constexpr int kTicketSlots = 64, kTicketSlotWords = 64;
struct TicketRing {
  int device;
  unsigned* counters;             // [kTicketSlots][kTicketSlotWords] device
  unsigned* finished;             // [kTicketSlots] device
  unsigned* done_dev;             // [kTicketSlots] device view of done_host
  volatile unsigned* done_host;   // [kTicketSlots] mapped, page-locked
  unsigned assigned[kTicketSlots];
  unsigned next;
};
struct TicketLease {
  unsigned *counters, *finished, *done;
  unsigned seq;
};
static std::atomic<unsigned> g_ticket_seq{0};
static std::mutex g_ticket_mutex;
static std::vector<TicketRing*> g_ticket_rings;

static int ticket_slots() {
  static const int n = [] {
    const char* e = getenv("CPF_TICKET_SLOTS");
    const int v = e ? atoi(e) : kTicketSlots;
    return v < 1 ? 1 : (v > kTicketSlots ? kTicketSlots : v);
  }();
  return n;
}

// *ok = false: every candidate slot is still in flight, the caller uses the static split
static int ticket_acquire(int device, TicketLease* lease, bool* ok) {
  std::lock_guard<std::mutex> lock(g_ticket_mutex);
  TicketRing* ring = nullptr;
  for (auto* e : g_ticket_rings)
    if (e->device == device) ring = e;
  if (!ring) {
    ring = new TicketRing();
    ring->device = device;
    ring->next = 0;
    memset(ring->assigned, 0, sizeof(ring->assigned));
    void* host = nullptr;
    CPF_CUDA(cudaMalloc(&ring->counters, (size_t)kTicketSlots * kTicketSlotWords * sizeof(unsigned)));
    CPF_CUDA(cudaMemset(ring->counters, 0, (size_t)kTicketSlots * kTicketSlotWords * sizeof(unsigned)));
    CPF_CUDA(cudaMalloc(&ring->finished, kTicketSlots * sizeof(unsigned)));
    CPF_CUDA(cudaMemset(ring->finished, 0, kTicketSlots * sizeof(unsigned)));
    CPF_CUDA(cudaStreamSynchronize(cudaStreamLegacy));      // the memsets run on the legacy stream, the kernels on non-blocking ones
    CPF_CUDA(cudaHostAlloc(&host, kTicketSlots * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(host, 0, kTicketSlots * sizeof(unsigned));
    ring->done_host = static_cast<volatile unsigned*>(host);
    CPF_CUDA(cudaHostGetDevicePointer((void**)&ring->done_dev, host, 0));
    g_ticket_rings.push_back(ring);
  }
  const unsigned slot = ring->next % (unsigned)ticket_slots();
  if (ring->assigned[slot] != 0 && ring->done_host[slot] != ring->assigned[slot]) {      // its last launch has not completed yet
    *ok = false;
    g_stat_fallback.fetch_add(1);
    return CPF_OK;
  }
  ring->next++;
  unsigned seq = g_ticket_seq.fetch_add(1u) + 1u;
  if (seq == 0) seq = g_ticket_seq.fetch_add(1u) + 1u;
  ring->assigned[slot] = seq;
  lease->counters = ring->counters + (size_t)slot * kTicketSlotWords;
  lease->finished = ring->finished + slot;
  lease->done = ring->done_dev + slot;
  lease->seq = seq;
  *ok = true;
  g_stat_dynamic.fetch_add(1);
  return CPF_OK;
}

template <int R1>
static int launch_pp(const FftlogArgs& a, const double2* tab, cudaStream_t stream) {
  typedef Geo<R1> G;
  constexpr int NG = 512 / G::T;
  size_t smem = (size_t)NG * G::SMEM_ELEMS * sizeof(double2);
  typedef void (*kern_t)(const FftlogArgs, const double2*);
  kern_t kern = fftlog_pp_kernel<R1, false>;
  // full window + 16-byte aligned rows: input rows staged by bulk copies (CPF_STREAM_TMA=0: direct loads)
  const char* tma_env = getenv("CPF_STREAM_TMA");
  const bool fullwin = a.n == G::N / 2 && a.in_left == G::N / 4 && !a.keep_padding;
  int dev = 0, sms = 0;
  CPF_CUDA(cudaGetDevice(&dev));
  CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = (a.pairs_per_p + NG - 1) / NG;
  if (grid > sms) grid = sms;
  FftlogArgs b = a;
  b.tickets = nullptr;
  if (fullwin && R1 < 16 && !(tma_env && tma_env[0] == '0') && ((uintptr_t)a.in % 16 == 0)) {
    kern = fftlog_pp_kernel<R1, true>;
    smem += (size_t)NG * G::N * sizeof(double);
    // ticket counters bring nothing here (r02u: 175-176 vs 176-178 M transforms/s at nk = 1024): the interleaved static split of
    // this kernel is already balanced; kept behind CPF_PP_DYNAMIC=1 (covered by the GPU tests through that variable)
    const char* dyn_env = getenv("CPF_PP_DYNAMIC");
    if ((dyn_env && dyn_env[0] == '1') && a.P <= kTicketSlotWords && a.pairs_per_p >= 8LL * grid * NG) {
      TicketLease lease;
      bool ok = false;
      CPF_TRY(ticket_acquire(dev, &lease, &ok));
      if (ok) {
        b.tickets = lease.counters; b.t_finished = lease.finished; b.t_done = lease.done; b.t_seq = lease.seq;
        kern = fftlog_pp_kernel<R1, true, true>;
      }
    }
  }
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const char* pdl_env = getenv("CPF_STREAM_PDL");
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_env && pdl_env[0] == '0') ? 0 : 1;
  CPF_CUDA(cudaLaunchKernelEx(&cfg, kern, b, tab));
  return CPF_OK;
}

static int launch_stream(const cpf_plan* pl, const FftlogArgs& a, cudaStream_t stream) {
  if (a.pairs_per_p > 2147483000LL) return fail(CPF_EUNSUPPORTED, "cpf_fftlog: batch too large for one launch");
  const bool fullwin = a.n == a.N / 2 && a.in_left == a.N / 4 && a.out_left == a.N / 4;
  StreamArgs s;
  s.in = a.in; s.out = a.out; s.pre = a.pre; s.post = a.post_re;
  s.in_row = a.in_has_P ? (long long)a.P * a.n : (long long)a.n;
  s.in_p = a.in_has_P ? a.n : 0;
  s.out_row = (long long)a.P * a.n_out;
  s.items = (long long)a.P * a.pairs_per_p;
  s.n = a.n; s.n_out = a.n_out; s.P = a.P; s.pairs_per_p = (int)a.pairs_per_p;
  s.odd_pair = (a.batch & 1) ? (int)(a.batch / 2) : -1;
  s.off_in = a.N / 4 - a.in_left; s.off_out = a.N / 4 - a.out_left;
  s.lines = (a.n * 8 + 127) / 128;
  s.dbg = nullptr;
  s.skew_ns = 0;
  // Phase offset inside a group: the two warps of a group that share an SM sub-partition leave every group barrier together and then want
  // the shared-memory pipe, and later the fp64 pipe, at the same moments.  Warps 4..7 of a group (the second warp of each sub-partition) idle
  // 250 ns after the SECOND group barrier of a pair, which takes them out of step (r5f-r5h: 78.2 -> 79.8 M transforms/s over 200 launches,
  // 77.0 -> 77.8 M over 20; 150-500 ns all give this, 550 ns and more lose; after the first barrier instead: +0.4 %).
  // CPF_STREAM_WSKEW2_NS / CPF_STREAM_WSKEW_NS (first barrier) / CPF_STREAM_WSKEW_INV (first barrier: the other warps) /
  // CPF_STREAM_WSKEW_MASK (which warps: warp index & mask) override.
  s.wskew_ns = 300;          // ... and warps 0..3 idle 300 ns after the FIRST group barrier (the other half: r5o/r5p, another +0.6 %)
  if (const char* e = getenv("CPF_STREAM_WSKEW_NS")) s.wskew_ns = atoi(e);
  s.wskew_mask = 4;
  if (const char* e = getenv("CPF_STREAM_WSKEW_MASK")) s.wskew_mask = atoi(e);
  s.wskew_inv = 1;
  if (const char* e = getenv("CPF_STREAM_WSKEW_INV")) s.wskew_inv = atoi(e);
  s.early = 1;
  if (const char* e = getenv("CPF_STREAM_EARLY")) s.early = atoi(e);
  s.wskew2_ns = 250;
  if (const char* e = getenv("CPF_STREAM_WSKEW2_NS")) s.wskew2_ns = atoi(e);
#ifdef CPF_LAB
  if (const char* e = getenv("CPF_STREAM_SKEW_NS")) s.skew_ns = atoi(e);
#endif
  int dev = 0, sms = 0;
  CPF_CUDA(cudaGetDevice(&dev));
  CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = (s.items + 1) / 2;
  if (grid > sms) grid = sms;
  typedef void (*kern_t)(const StreamArgs, const double*, const double2*, const double2*);
  kern_t kern = fullwin ? (kern_t)fftlog_stream_kernel<true> : (kern_t)fftlog_stream_kernel<false>;
  // staged variant: bulk copies need 16-byte aligned rows; CPF_STREAM_TMA=0 selects the direct-load kernel (A/B runs)
  int smem_bytes = ST_SMEM_BYTES;
  const char* tma_env = getenv("CPF_STREAM_TMA");
  const bool want_tma = !tma_env || tma_env[0] != '0';
  s.tickets = s.t_finished = s.t_done = nullptr;
  s.t_seq = 0;
  if (fullwin && want_tma && ((uintptr_t)a.in % 16 == 0) && (s.in_row % 2 == 0) && (s.in_p % 2 == 0)) {
    kern = (kern_t)fftlog_stream_kernel<true, 0, true>;
    smem_bytes = ST_SMEM_BYTES_TMA;
    // dynamic scheduling when every CTA owns one plan row and has a queue worth drawing from (CPF_STREAM_DYNAMIC=0: static split)
    const char* dyn_env = getenv("CPF_STREAM_DYNAMIC");
    const long long ctas_per_row = (grid + a.P - 1) / a.P;
    if (!(dyn_env && dyn_env[0] == '0') && grid >= a.P && a.P <= kTicketSlotWords && a.pairs_per_p >= 8 * ctas_per_row) {
      TicketLease lease;
      bool ok = false;
      CPF_TRY(ticket_acquire(dev, &lease, &ok));
      if (ok) {
        s.tickets = lease.counters; s.t_finished = lease.finished; s.t_done = lease.done; s.t_seq = lease.seq;
        kern = (kern_t)fftlog_stream_kernel<true, 0, true, true>;
      }
    }
  }
#ifdef CPF_LAB
  if (const char* e = getenv("CPF_STREAM_ABL")) {
    switch (atoi(e)) {
      case 1: kern = fftlog_stream_kernel<true, 1>; break;
      case 4: kern = fftlog_stream_kernel<true, 4>; break;
      case 5: kern = fftlog_stream_kernel<true, 5>; break;
      case 13: kern = fftlog_stream_kernel<true, 13>; break;
      case 15: kern = fftlog_stream_kernel<true, 15>; break;
      case 16: kern = fftlog_stream_kernel<true, 16>; break;
      case 32: kern = fftlog_stream_kernel<true, 32>; break;
      case 48: kern = fftlog_stream_kernel<true, 48>; break;
      default: break;
    }
  }
#endif
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
#ifdef CPF_LAB
  static long long* d_dbg = nullptr;
  if (getenv("CPF_STREAM_DBG")) {
    if (!d_dbg) CPF_CUDA(cudaMalloc(&d_dbg, 8 * 256 * sizeof(long long)));
    CPF_CUDA(cudaMemsetAsync(d_dbg, 0, 8 * 256 * sizeof(long long), stream));
    s.dbg = d_dbg;
  }
#endif
  // launched with programmatic stream serialisation: back-to-back calls on a stream overlap this kernel's prologue with the
  // tail of the previous one (the kernel executes griddepcontrol.wait before it touches rows); CPF_STREAM_PDL=0 turns it off
  const char* pdl_env = getenv("CPF_STREAM_PDL");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_env && pdl_env[0] == '0') ? 0 : 1;
  CPF_CUDA(cudaLaunchKernelEx(&cfg, kern, s, (const double*)pl->fast->d_st_tw, (const double2*)pl->d_st_ut, (const double2*)pl->fast->d_m256));
#ifdef CPF_LAB
  if (s.dbg && getenv("CPF_STREAM_DBG")[0] == '2') {     // print the time line of this launch (synchronises)
    std::vector<long long> h(8 * 256);
    CPF_CUDA(cudaStreamSynchronize(stream));
    CPF_CUDA(cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long t0 = h[0], tend = 0;
    for (long long b = 0; b < grid; ++b) { if (h[8 * b] < t0) t0 = h[8 * b]; if (h[8 * b + 6] > tend) tend = h[8 * b + 6]; }
    printf("stream kernel time line (us after the first CTA start; grid %lld, span %.2f us)\n", grid, (tend - t0) * 1e-3);
    printf("  cta   start  tw-done  p-tables  1st-pair  seg-end  loop-end  exit\n");
    for (long long b = 0; b < grid; b += (grid > 16 ? grid / 12 : 1)) {
      printf("  %3lld", b);
      for (int k = 0; k < 7; ++k) printf(" %8.2f", (h[8 * b + k] - t0) * 1e-3);
      printf("\n");
    }
    double mx[7] = {0}, mn[7];
    for (int k = 0; k < 7; ++k) mn[k] = 1e30;
    for (long long b = 0; b < grid; ++b)
      for (int k = 0; k < 7; ++k) { const double v = (h[8 * b + k] - t0) * 1e-3; if (v > mx[k]) mx[k] = v; if (v < mn[k]) mn[k] = v; }
    printf("  min"); for (int k = 0; k < 7; ++k) printf(" %8.2f", mn[k]); printf("\n  max"); for (int k = 0; k < 7; ++k) printf(" %8.2f", mx[k]); printf("\n");
  }
#endif
  CPF_CUDA(cudaGetLastError());
  return CPF_OK;
}

// ping-pong tables of an N = 4096 plan, built the first time that kernel is asked for
static int ensure_pp_tables(const cpf_plan* pl) {
  std::lock_guard<std::mutex> lock(pl->pp_mutex);
  if (pl->d_pp) return CPF_OK;
  if (pl->h_uhs.empty()) return fail(CPF_EUNSUPPORTED, "cpf_fftlog: this plan has no ping-pong tables");
  std::vector<double2> tab;
  build_pp_tables(pl->N, pl->fast_R1, pl->P, pl->h_pre_win.data(), pl->h_uhs, pl->h_post_win.data(), pl->fast->pp_tw, tab);
  double2* d = nullptr;
  CPF_TRY(upload((void**)&d, tab.data(), tab.size() * sizeof(double2)));
  pl->d_pp = d;
  return CPF_OK;
}

// records of the persistent N = 8192 kernel, built the first time it is asked for
static int ensure_pp8k_tables(const cpf_plan* pl) {
  std::lock_guard<std::mutex> lock(pl->pp_mutex);
  if (pl->d_pp8k) return CPF_OK;
  if (pl->h_uhs.empty()) return fail(CPF_EUNSUPPORTED, "cpf_fftlog: this plan has no tables for the persistent N = 8192 kernel");
  std::vector<double2> tab;
  build_pp8k_tables(pl->P, pl->h_pre_win.data(), pl->h_uhs, pl->h_post_win.data(), pl->fast->pp_tw, tab);
  double2* d = nullptr;
  CPF_TRY(upload((void**)&d, tab.data(), tab.size() * sizeof(double2)));
  pl->d_pp8k = d;
  return CPF_OK;
}

static int launch_pp8k(const cpf_plan* pl, const FftlogArgs& a, cudaStream_t stream) {
  int dev = 0, sms = 0;
  CPF_CUDA(cudaGetDevice(&dev));
  CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = (long long)a.P * a.pairs_per_p;        // a CTA serves one plan row when there are at least P of them
  if (grid > sms) grid = sms;
  typedef void (*kern_t)(const FftlogArgs, const double2*, const double2*, const double2*);
  kern_t kern = fftlog_pp8k_kernel<false>;
  int smem = P8_SMEM_BYTES;
  // full window + 16-byte aligned rows: the rows of a pair are staged by bulk copies (CPF_STREAM_TMA=0: direct loads)
  const char* tma_env = getenv("CPF_STREAM_TMA");
  const bool fullwin = a.n == a.N / 2 && a.in_left == a.N / 4 && !a.keep_padding;
  if (fullwin && !(tma_env && tma_env[0] == '0') && ((uintptr_t)a.in % 16 == 0)) {
    kern = fftlog_pp8k_kernel<true>;
    smem = P8_SMEM_BYTES_TMA;
  }
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const char* pdl_env = getenv("CPF_STREAM_PDL");
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_env && pdl_env[0] == '0') ? 0 : 1;
  FftlogArgs b = a;
  b.tickets = nullptr;
  // phase offset between the two chains of a pair: chain 1 idles this long after the barrier that follows the input, so that its
  // shared-memory phases fall on chain 0's fp64 phases instead of on its shared-memory phases (r5b/r5c: 0 ns 26.07, 400 ns 26.38, 800 ns
  // 26.72, 1600 ns 26.61, 2400 ns 25.16 M transforms/s); CPF_PP8K_SKEW_NS overrides
  b.ahead = 800;
  if (const char* e = getenv("CPF_PP8K_SKEW_NS")) b.ahead = atoi(e);
  CPF_CUDA(cudaLaunchKernelEx(&cfg, kern, b, (const double2*)pl->d_pp8k, pl->d_tw2_256, pl->d_tw8192));
  return CPF_OK;
}

// Kernel choice for the default call (zero padding, cropped output, real post-factor).  CPF_FFTLOG_KERNEL = fast | pp |
// stream forces one family (where the plan has its tables); otherwise large launches go to the persistent kernels
// (stream for N = 4096, ping-pong for N = 2048 / 1024) and small ones to the per-pair kernel, which has no start-up cost.
enum { K_AUTO = -1, K_FAST = 0, K_PP = 1, K_STREAM = 2 };
static int kernel_choice() {
  const char* e = getenv("CPF_FFTLOG_KERNEL");
  if (!e) return K_AUTO;
  if (e[0] == 'f') return K_FAST;
  if (e[0] == 'p') return K_PP;
  if (e[0] == 's') return K_STREAM;
  return K_AUTO;
}

static int generic_threads(int N) {
  int t = N / 2;
  if (t < 32) t = 32;
  if (t > 512) t = 512;
  return t;
}

// one launch over `batch` device-resident rows
static int launch_fftlog(const cpf_plan* pl, FftlogArgs a, bool pruned, cudaStream_t stream) {
  if (a.batch <= 0) return CPF_OK;
  a.pairs_per_p = (a.batch + 1) / 2;
  const long long nblocks = (long long)pl->P * a.pairs_per_p;
  if (nblocks > 2147483647LL) return fail(CPF_EUNSUPPORTED, "cpf_fftlog: batch too large for one launch");
  if (pl->fast_R1) {
    a.ut = pruned ? pl->d_uts : pl->d_ut;
    a.tw1 = pl->fast->d_tw1;
    a.tw2 = pl->fast->d_tw2;
    if (pruned && (pl->d_pp || pl->d_st_ut)) {     // (these tables exist only for prunable windows and a real post-factor)
      const int choice = kernel_choice();
      int dev = 0, sms = 0;
      CPF_CUDA(cudaGetDevice(&dev));
      CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const bool big = nblocks >= 16LL * sms;     // >= 8 pairs per 256-thread group: start-up and tail are amortised
      if (pl->d_st_ut && (choice == K_STREAM || (choice == K_AUTO && big))) return launch_stream(pl, a, stream);
      if (choice == K_PP || (choice == K_AUTO && big && !pl->d_st_ut)) {
        if (!pl->d_pp) CPF_TRY(ensure_pp_tables(pl));
        switch (pl->fast_R1) {
          case 16: return launch_pp<16>(a, pl->d_pp, stream);
          case 8: return launch_pp<8>(a, pl->d_pp, stream);
          default: return launch_pp<4>(a, pl->d_pp, stream);
        }
      }
    }
    switch (pl->fast_R1) {
      case 16: return launch_fast_r<16>(a, pruned, pl->post_complex, nblocks, stream);
      case 8: return launch_fast_r<8>(a, pruned, pl->post_complex, nblocks, stream);
      default: return launch_fast_r<4>(a, pruned, pl->post_complex, nblocks, stream);
    }
  }
  if (pl->split2) {
    a.ut = pruned ? pl->d_uts : pl->d_ut;
    a.tw1 = pl->fast->d_tw1;
    a.tw2 = pl->fast->d_tw2;
    if (pruned && !pl->post_complex && !pl->h_uhs.empty()) {       // the default call: large launches go to the persistent kernel
      const int choice = kernel_choice();
      int dev = 0, sms = 0;
      CPF_CUDA(cudaGetDevice(&dev));
      CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      if (choice == K_PP || (choice == K_AUTO && a.pairs_per_p >= 4LL * sms)) {
        CPF_TRY(ensure_pp8k_tables(pl));
        return launch_pp8k(pl, a, stream);
      }
    }
    const size_t smem2 = (size_t)Geo<16>::SMEM_ELEMS * sizeof(double2);
    typedef void (*kern2_t)(const FftlogArgs, const double2*);
    kern2_t kern = pruned ? (pl->post_complex ? (kern2_t)fftlog_split2_kernel<true, true> : (kern2_t)fftlog_split2_kernel<true, false>)
                          : (pl->post_complex ? (kern2_t)fftlog_split2_kernel<false, true> : (kern2_t)fftlog_split2_kernel<false, false>);
    CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    if (2 * nblocks > 2147483647LL) return fail(CPF_EUNSUPPORTED, "cpf_fftlog: batch too large for one launch");
    {
      int dev = 0, sms = 0;
      CPF_CUDA(cudaGetDevice(&dev));
      CPF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      a.ahead = sms;                                    // two CTAs per SM = `sms` clusters per wave
    }
    kern<<<(unsigned)(2 * nblocks), 256, smem2, stream>>>(a, pl->d_tw8192);      // clusters of two CTAs (__cluster_dims__)
    CPF_CUDA(cudaGetLastError());
    return CPF_OK;
  }
  a.ut = pl->d_ut;
  a.tw1 = pl->d_tw;
  a.tw2 = nullptr;
  const size_t smem = (size_t)pl->N * sizeof(double2);
  if (pl->post_complex) {
    CPF_CUDA(cudaFuncSetAttribute(fftlog_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fftlog_generic_kernel<true><<<(unsigned)nblocks, generic_threads(pl->N), smem, stream>>>(a);
  } else {
    CPF_CUDA(cudaFuncSetAttribute(fftlog_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fftlog_generic_kernel<false><<<(unsigned)nblocks, generic_threads(pl->N), smem, stream>>>(a);
  }
  CPF_CUDA(cudaGetLastError());
  return CPF_OK;
}

// How host-pointer calls move their data: CPF_HOST_PATH = staged (default) | zerocopy | mixed.
//   staged   : chunks are copied in, processed and copied back on rotating streams (H2D / kernel / D2H overlap);
//   zerocopy : when both buffers are page-locked, the kernel reads the input rows and writes the output rows straight
//              over PCIe (unified addressing: one launch, both directions busy from the first to the last row).
static bool host_zero_copy_enabled() {
  const char* e = getenv("CPF_HOST_PATH");
  return e && e[0] == 'z';
}
//   mixed    : inputs are staged as above, results are written by the kernels straight into the (page-locked) host buffer:
//              posted PCIe writes from the SMs, no device-to-host copies (and no dependency chain between the two copy engines).
static bool host_direct_out_enabled() {
  const char* e = getenv("CPF_HOST_PATH");
  return e && e[0] == 'm';
}

static bool pinned_device_pointer(const void* host, void** dev) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}

// Pageable host input (an ordinary numpy array): cudaMemcpyAsync stages such memory through the driver's own bounce buffer on the calling
// thread at ~10 GB/s and blocks while it does.  Instead a few persistent worker threads copy every chunk, in slices, into page-locked
// buffers of the staging pool one chunk ahead of its H2D copy (CPF_HOST_THREADS, default 8, 0 = let the driver do it).
static bool is_pageable_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}
struct HostCopyPool {
  struct Job { char* dst; const char* src; size_t bytes; std::atomic<int>* pending; };
  std::mutex m;
  std::condition_variable cv, done;
  std::deque<Job> q;
  int nthreads = 0;
  void worker() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [&] { return !q.empty(); });
        j = q.front();
        q.pop_front();
      }
      memcpy(j.dst, j.src, j.bytes);
      if (j.pending->fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> lock(m);
        done.notify_all();
      }
    }
  }
  // copy `bytes` in up to nthreads slices; *pending counts the slices still running
  void submit(char* dst, const char* src, size_t bytes, std::atomic<int>* pending) {
    const size_t slice = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    int n = 0;
    for (size_t o = 0; o < bytes; o += slice) ++n;
    pending->store(n);
    std::lock_guard<std::mutex> lock(m);
    for (size_t o = 0; o < bytes; o += slice) q.push_back({dst + o, src + o, bytes - o < slice ? bytes - o : slice, pending});
    cv.notify_all();
  }
  void wait(std::atomic<int>* pending) {
    std::unique_lock<std::mutex> lock(m);
    done.wait(lock, [&] { return pending->load() == 0; });
  }
};
static HostCopyPool* host_copy_pool() {          // created on first use, lives until the process exits (the workers are detached)
  static std::mutex mu;
  static HostCopyPool* pool = nullptr;
  static long owner = -1;                         // threads do not survive fork(): a child process makes its own pool
  std::lock_guard<std::mutex> lock(mu);
  const long me = (long)getpid();
  if (owner != me) {
    owner = me;
    pool = nullptr;
    const char* e = getenv("CPF_HOST_THREADS");
    int n = e ? atoi(e) : 8;
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && n > hw) n = hw;
    if (n > 0) {
      HostCopyPool* p = new HostCopyPool();
      p->nthreads = n;
      for (int i = 0; i < n; ++i) std::thread([p] { p->worker(); }).detach();
      pool = p;
    }
  }
  return pool;
}

// Chunk sizes (rows) of a staged host-pointer call: small chunks at both ends so that the pipeline fills and drains
// quickly (the first H2D and the last D2H are not overlapped with anything), `cap`-sized chunks in the middle where
// per-chunk overheads matter.  Every chunk but the last has an even number of rows.
static std::vector<long long> chunk_schedule(long long rows, long long small, long long cap) {
  // scheduled in row pairs; an odd row count shortens the last chunk by one row
  const long long sp = small / 2 > 0 ? small / 2 : 1, cp = cap / 2 > sp ? cap / 2 : sp;
  long long left = (rows + 1) / 2;
  std::vector<long long> front, back;
  for (long long c = sp; c < cp && left > 0; c *= 2) {
    const long long f = c < left ? c : left;
    front.push_back(f);
    left -= f;
    const long long b = c < left ? c : left;
    if (b > 0) back.push_back(b);
    left -= b;
  }
  for (; left > 0; left -= (cp < left ? cp : left)) front.push_back(cp < left ? cp : left);
  for (size_t i = back.size(); i-- > 0;) front.push_back(back[i]);
  for (auto& v : front) v *= 2;
  if (rows & 1) front.back() -= 1;
  return front;
}

// Runs `body(chunk_first_row, chunk_rows, d_in_chunk, d_out_chunk, stream)` over the batch.  Device-resident calls
// are one chunk on the caller's stream.  Host-pointer calls are split into chunks that are copied in, processed and
// copied back on rotating internal streams so that H2D, compute and D2H overlap; the call returns when all chunks
// are back in the caller's buffer.
template <typename Body>
static int run_staged(int device, const double* in, size_t in_row_doubles, double* out, size_t out_row_doubles,
                      long long rows, bool in_dev, bool out_dev, cudaStream_t user_stream, Body body) {
  if (rows <= 0) return CPF_OK;
  if (in_dev && out_dev) return body(0LL, rows, in, out, user_stream);
  if (!in_dev && !out_dev && host_zero_copy_enabled()) {
    void *zin = nullptr, *zout = nullptr;
    if (pinned_device_pointer(in, &zin) && pinned_device_pointer(out, &zout)) {
      CPF_TRY(body(0LL, rows, (const double*)zin, (double*)zout, user_stream));
      CPF_CUDA(cudaStreamSynchronize(user_stream));
      return CPF_OK;
    }
  }
  StagePool* sp = nullptr;
  CPF_TRY(stage_pool(device, &sp));
  std::lock_guard<std::mutex> lock(sp->busy);
  const size_t row_bytes = (in_row_doubles > out_row_doubles ? in_row_doubles : out_row_doubles) * sizeof(double);
  // Small calls (a single transform, the reference's configs[0]): nothing to overlap, so no chunks, events or copy streams: copy in,
  // run, copy out on one stream: a third of the API calls of the pipelined path (83 -> ~50 us for one nk = 1024 transform from Python).
  if ((size_t)rows * row_bytes <= kSmallCallBytes && !host_direct_out_enabled()) {
    ScratchBuf din1, dout1;
    const double* d_in = in;
    double* d_out = out;
    const size_t in_bytes = (size_t)rows * in_row_doubles * sizeof(double), out_bytes = (size_t)rows * out_row_doubles * sizeof(double);
    if (!in_dev && !out_dev) {
      // both sides on the host: the kernel reads and writes two mapped page-locked buffers over PCIe, so the call is two small memcpys, one
      // launch and one synchronisation (no copy-engine round trips: 34 -> ~20 us for one nk = 1024 transform)
      if (!sp->h_small[0]) {
        for (int i = 0; i < 2; ++i) {
          CPF_CUDA(cudaHostAlloc(&sp->h_small[i], kSmallCallBytes, cudaHostAllocMapped | cudaHostAllocPortable));
          CPF_CUDA(cudaHostGetDevicePointer(&sp->d_small[i], sp->h_small[i], 0));
        }
      }
      memcpy(sp->h_small[0], in, in_bytes);
      CPF_TRY(body(0LL, rows, (const double*)sp->d_small[0], (double*)sp->d_small[1], sp->comp));
      CPF_CUDA(cudaStreamSynchronize(sp->comp));
      memcpy(out, sp->h_small[1], out_bytes);
      return CPF_OK;
    }
    if (!in_dev) {
      CPF_CUDA(din1.alloc(in_bytes, sp->comp));
      CPF_CUDA(cudaMemcpyAsync(din1.p, in, in_bytes, cudaMemcpyHostToDevice, sp->comp));
      d_in = (const double*)din1.p;
    }
    if (!out_dev) {
      CPF_CUDA(dout1.alloc(out_bytes, sp->comp));
      d_out = (double*)dout1.p;
    }
    if (in_dev || out_dev) {                      // one side lives on the caller's stream: order the internal stream after / before it
      CPF_CUDA(cudaEventRecord(sp->start, user_stream));
      CPF_CUDA(cudaStreamWaitEvent(sp->comp, sp->start, 0));
    }
    CPF_TRY(body(0LL, rows, d_in, d_out, sp->comp));
    if (!out_dev) CPF_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, sp->comp));
    CPF_CUDA(cudaStreamSynchronize(sp->comp));
    return CPF_OK;
  }
  // tuning knobs (bytes / count): CPF_STAGE_CAP_KB (largest chunk), CPF_STAGE_SMALL_KB (first and last chunk), CPF_STAGE_NBUF
  size_t cap_bytes = 16u << 20, small_bytes = 1u << 20;
  int max_buf = 4;
  if (const char* e = getenv("CPF_STAGE_CAP_KB")) cap_bytes = (size_t)atoll(e) << 10;
  if (const char* e = getenv("CPF_STAGE_SMALL_KB")) small_bytes = (size_t)atoll(e) << 10;
  if (const char* e = getenv("CPF_STAGE_NBUF")) max_buf = atoi(e);
  if (max_buf < 1) max_buf = 1;
  if (max_buf > kNumStage) max_buf = kNumStage;
  long long cap = (long long)(cap_bytes / (row_bytes ? row_bytes : 1));
  cap = cap < 2 ? 2 : (cap & ~1LL);                 // even, so that row pairs never straddle chunks
  long long small = (long long)(small_bytes / (row_bytes ? row_bytes : 1));
  small = small < 2 ? 2 : (small & ~1LL);
  if (small > cap) small = cap;                     // chunk_schedule never emits more than max(small, cap) rows: keep that <= buf_rows
  const std::vector<long long> sched = chunk_schedule(rows, small, cap);
  void* zout = nullptr;
  const bool direct_out = !out_dev && host_direct_out_enabled() && pinned_device_pointer(out, &zout);
  if (direct_out) out_dev = true;          // from here on the output is "a device pointer": the mapped host buffer
  double* const out_base = direct_out ? (double*)zout : out;
  const int nbuf = sched.size() < (size_t)max_buf ? (int)sched.size() : max_buf;
  const long long buf_rows = rows < cap ? rows : cap;
  ScratchBuf din[kNumStage], dout[kNumStage];
  CPF_CUDA(cudaEventRecord(sp->start, user_stream));
  CPF_CUDA(cudaStreamWaitEvent(sp->h2d, sp->start, 0));
  CPF_CUDA(cudaStreamWaitEvent(sp->comp, sp->start, 0));
  CPF_CUDA(cudaStreamWaitEvent(sp->d2h, sp->start, 0));
  for (int i = 0; i < nbuf; ++i) {
    if (!in_dev) CPF_CUDA(din[i].alloc((size_t)buf_rows * in_row_doubles * sizeof(double), sp->comp));
    if (!out_dev) CPF_CUDA(dout[i].alloc((size_t)buf_rows * out_row_doubles * sizeof(double), sp->comp));
  }
  // the copy streams may touch the buffers only after the (stream-ordered) allocations on the compute stream
  CPF_CUDA(cudaEventRecord(sp->start, sp->comp));
  CPF_CUDA(cudaStreamWaitEvent(sp->h2d, sp->start, 0));
  CPF_CUDA(cudaStreamWaitEvent(sp->d2h, sp->start, 0));
  int rc = CPF_OK;
  long long first = 0;
  // pageable input: worker threads fill page-locked bounce buffers one chunk ahead (see HostCopyPool)
  HostCopyPool* hpool = (!in_dev && rows * (long long)in_row_doubles * 8 >= (4LL << 20) && is_pageable_host(in)) ? host_copy_pool() : nullptr;
  std::atomic<int> bounce_pending[kNumStage];
  std::vector<long long> chunk_first(sched.size() + 1, 0);
  for (size_t c = 0; c < sched.size(); ++c) chunk_first[c + 1] = chunk_first[c] + sched[c];
  if (hpool) {
    const size_t need = (size_t)buf_rows * in_row_doubles * sizeof(double);
    if (sp->h_bounce_bytes < need) {
      for (int i = 0; i < kNumStage; ++i) { if (sp->h_bounce[i]) cudaFreeHost(sp->h_bounce[i]); sp->h_bounce[i] = nullptr; }
      sp->h_bounce_bytes = 0;
      bool ok = true;
      for (int i = 0; i < max_buf && ok; ++i) ok = cudaHostAlloc(&sp->h_bounce[i], need, cudaHostAllocPortable) == cudaSuccess;
      if (ok) sp->h_bounce_bytes = need;
      else { cudaGetLastError(); hpool = nullptr; }        // no page-locked memory to spare: the driver stages the copies
    }
  }
  auto bounce_fill = [&](const size_t c) -> cudaError_t {   // start copying chunk c into its bounce buffer
    const int i = (int)(c % nbuf);
    if (c >= (size_t)nbuf) {                                 // the H2D copy of chunk c - nbuf has left the buffer
      const cudaError_t e = cudaEventSynchronize(sp->in_ready[i]);
      if (e != cudaSuccess) return e;
    }
    hpool->submit((char*)sp->h_bounce[i], (const char*)(in + (size_t)chunk_first[c] * in_row_doubles), (size_t)sched[c] * in_row_doubles * sizeof(double),
                  &bounce_pending[i]);
    return cudaSuccess;
  };
  if (hpool) {
    const cudaError_t e = bounce_fill(0);
    if (e != cudaSuccess) rc = fail(CPF_ECUDA, "event sync: %s", cudaGetErrorString(e));
  }
  for (size_t c = 0; c < sched.size() && rc == CPF_OK; ++c) {
    const int i = (int)(c % nbuf);
    const bool reused = c >= (size_t)nbuf;
    const long long cnt = sched[c];
    const double* src = in + (size_t)first * in_row_doubles;
    if (hpool) {
      // the next chunk goes into its bounce buffer while this one crosses the link (one buffer only: this chunk, now)
      const size_t ahead = nbuf > 1 ? c + 1 : c;
      if (ahead > 0 && ahead < sched.size()) {
        const cudaError_t e = bounce_fill(ahead);
        if (e != cudaSuccess) { rc = fail(CPF_ECUDA, "event sync: %s", cudaGetErrorString(e)); break; }
      }
      hpool->wait(&bounce_pending[i]);
      src = (const double*)sp->h_bounce[i];
    }
    double* dst = out_base + (size_t)first * out_row_doubles;
    const double* d_in = src;
    double* d_out = dst;
    cudaError_t e = cudaSuccess;
    if (!in_dev) {
      d_in = (const double*)din[i].p;
      if (reused) e = cudaStreamWaitEvent(sp->h2d, sp->k_done[i], 0);          // the kernel that read buffer i is done
      if (e == cudaSuccess) e = cudaMemcpyAsync(din[i].p, src, (size_t)cnt * in_row_doubles * sizeof(double), cudaMemcpyHostToDevice, sp->h2d);
      if (e == cudaSuccess) e = cudaEventRecord(sp->in_ready[i], sp->h2d);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(sp->comp, sp->in_ready[i], 0);
      if (e != cudaSuccess) { rc = fail(CPF_ECUDA, "H2D copy: %s", cudaGetErrorString(e)); break; }
    }
    if (!out_dev) {
      d_out = (double*)dout[i].p;
      if (reused) e = cudaStreamWaitEvent(sp->comp, sp->out_free[i], 0);       // the D2H copy out of buffer i is done
      if (e != cudaSuccess) { rc = fail(CPF_ECUDA, "stream wait: %s", cudaGetErrorString(e)); break; }
    }
    rc = body(first, cnt, d_in, d_out, sp->comp);
    if (rc != CPF_OK) break;
    e = cudaEventRecord(sp->k_done[i], sp->comp);
    if (e == cudaSuccess && !out_dev) {
      e = cudaStreamWaitEvent(sp->d2h, sp->k_done[i], 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dst, d_out, (size_t)cnt * out_row_doubles * sizeof(double), cudaMemcpyDeviceToHost, sp->d2h);
      if (e == cudaSuccess) e = cudaEventRecord(sp->out_free[i], sp->d2h);
    }
    if (e != cudaSuccess) { rc = fail(CPF_ECUDA, "D2H copy: %s", cudaGetErrorString(e)); break; }
    first += cnt;
  }
  for (cudaStream_t st : {sp->h2d, sp->comp, sp->d2h}) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && rc == CPF_OK) rc = fail(CPF_ECUDA, "stream sync: %s", cudaGetErrorString(e));
  }
  return rc;
}

}  // namespace cpf

extern "C" {

int cpf_fftlog(const cpf_plan* pl, const double* in, int64_t batch, int in_has_P, int ex_l_mode, double ex_l_val,
               int ex_r_mode, double ex_r_val, int keep_padding, double* out, int in_on_device, int out_on_device,
               void* stream) {
  if (!pl) return fail(CPF_EINVAL, "cpf_fftlog: null plan");
  if (batch < 0) return fail(CPF_EINVAL, "cpf_fftlog: negative batch");
  if (batch == 0) return CPF_OK;
  if (!in || !out) return fail(CPF_EINVAL, "cpf_fftlog: null buffer");
  for (int m : {ex_l_mode, ex_r_mode})
    if (m != CPF_EXTRAP_CONST && m != CPF_EXTRAP_EDGE && m != CPF_EXTRAP_LOG) return fail(CPF_EINVAL, "cpf_fftlog: unknown extrapolation mode %d", m);
  if ((ex_l_mode == CPF_EXTRAP_LOG || ex_r_mode == CPF_EXTRAP_LOG) && pl->n < 2)
    return fail(CPF_EINVAL, "cpf_fftlog: 'log' extrapolation needs at least two samples");
  DeviceGuard guard(pl->device);
  if (guard.err != cudaSuccess) return fail(CPF_ECUDA, "cudaSetDevice(%d): %s", pl->device, cudaGetErrorString(guard.err));

  FftlogArgs a;
  a.tickets = a.t_finished = a.t_done = nullptr;
  a.t_seq = 0;
  a.ahead = 0;
  a.pre = pl->d_pre;
  a.post_re = pl->d_post_re;
  a.post_im = pl->d_post_im;
  a.P = pl->P; a.in_has_P = in_has_P ? 1 : 0; a.n = pl->n; a.N = pl->N; a.log2N = pl->log2N;
  a.in_left = pl->in_left; a.out_left = pl->out_left;
  a.keep_padding = keep_padding ? 1 : 0;
  a.n_out = keep_padding ? pl->N : pl->n;
  a.ex_l_mode = ex_l_mode; a.ex_r_mode = ex_r_mode; a.ex_l_val = ex_l_val; a.ex_r_val = ex_r_val;
  const bool pruned = use_pruned(pl, ex_l_mode, ex_l_val, ex_r_mode, ex_r_val, keep_padding);
  const size_t in_row = (size_t)pl->n * (in_has_P ? pl->P : 1);
  const size_t out_row = (size_t)a.n_out * pl->P * (pl->post_complex ? 2 : 1);
  return run_staged(pl->device, in, in_row, out, out_row, batch, in_on_device != 0, out_on_device != 0, (cudaStream_t)stream,
                    [&](long long, long long cnt, const double* d_in, double* d_out, cudaStream_t s) {
                      FftlogArgs c = a;
                      c.in = d_in; c.out = d_out; c.batch = cnt;
                      return launch_fftlog(pl, c, pruned, s);
                    });
}

static int check_fft_size(const char* who, int size, int device) {
  if (!is_pow2(size) || size < 2) return fail(CPF_EINVAL, "%s: size=%d must be a power of two >= 2", who, size);
  if (size > CPF_MAX_N) return fail(CPF_EUNSUPPORTED, "%s: size=%d exceeds CPF_MAX_N=%d", who, size, CPF_MAX_N);
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "%s: device %d out of range (%d visible)", who, device, ndev);
  return CPF_OK;
}

int cpf_rfft(int size, const double* in, int64_t rows, double* out, int in_on_device, int out_on_device, int device,
             void* stream) {
  CPF_TRY(check_fft_size("cpf_rfft", size, device));
  if (rows <= 0) return rows < 0 ? fail(CPF_EINVAL, "cpf_rfft: negative rows") : CPF_OK;
  if (!in || !out) return fail(CPF_EINVAL, "cpf_rfft: null buffer");
  DeviceGuard guard(device);
  double2* tw = nullptr;
  CPF_TRY(generic_twiddles(device, size, &tw));
  const size_t smem = (size_t)size * sizeof(double2);
  CPF_CUDA(cudaFuncSetAttribute(rfft_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int l2 = ilog2(size);
  return run_staged(device, in, (size_t)size, out, (size_t)(size / 2 + 1) * 2, rows, in_on_device != 0, out_on_device != 0,
                    (cudaStream_t)stream, [&](long long, long long cnt, const double* d_in, double* d_out, cudaStream_t s) {
                      rfft_pair_kernel<<<(unsigned)((cnt + 1) / 2), generic_threads(size), smem, s>>>(d_in, (double2*)d_out, cnt, size, l2, tw);
                      CPF_CUDA(cudaGetLastError());
                      return (int)CPF_OK;
                    });
}

int cpf_irfft_conj(int size, const double* in, int64_t rows, double* out, int in_on_device, int out_on_device,
                   int device, void* stream) {
  CPF_TRY(check_fft_size("cpf_irfft_conj", size, device));
  if (rows <= 0) return rows < 0 ? fail(CPF_EINVAL, "cpf_irfft_conj: negative rows") : CPF_OK;
  if (!in || !out) return fail(CPF_EINVAL, "cpf_irfft_conj: null buffer");
  DeviceGuard guard(device);
  double2* tw = nullptr;
  CPF_TRY(generic_twiddles(device, size, &tw));
  const size_t smem = (size_t)size * sizeof(double2);
  CPF_CUDA(cudaFuncSetAttribute(irfft_conj_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int l2 = ilog2(size);
  return run_staged(device, in, (size_t)(size / 2 + 1) * 2, out, (size_t)size, rows, in_on_device != 0, out_on_device != 0,
                    (cudaStream_t)stream, [&](long long, long long cnt, const double* d_in, double* d_out, cudaStream_t s) {
                      irfft_conj_pair_kernel<<<(unsigned)((cnt + 1) / 2), generic_threads(size), smem, s>>>((const double2*)d_in, d_out, cnt, size, l2, tw);
                      CPF_CUDA(cudaGetLastError());
                      return (int)CPF_OK;
                    });
}

}  // extern "C"
