// cpf_fftlog_stream.cuh — persistent "stream" FFTLog kernel (sm_100a) for the default call at N = 4096 (zero padding,
// cropped output, real post-factor: cosmoprimo/fftlog.py:198-241 with extrap=0, keep_padding=False).  Included by
// cpf_fftlog.cu after cpf_fftlog_pp.cuh (tensor-memory helpers).
//
// Why (profiles/r01e_summary.md): in the ping-pong kernel the fp64 pipe is 62 % busy while `math_pipe_throttle` is
// the top stall: its two lock-stepped 256-thread groups fall into step, so the fp64 phases of both collide and the
// shared-memory phases of both collide.  Here the data flow of cpf_stream_core.h leaves only two group barriers per
// pair of rows (the two exchanges that really cross warps); everything in between — P2, P3, kernel multiply, P1',
// P2' and, across pairs, P3', store, load, P1 — runs per warp, so the 16 warps of an SM drift apart and keep both the
// fp64 pipe and the shared-memory pipe fed.
//
//   * one 512-thread CTA per SM, two groups of 256 threads, each transforming one pair of rows at a time; a CTA owns
//     a contiguous range of (plan row, pair) items so that it normally loads one plan row's tables once;
//   * thread-private tables (P1, P2, P1' twiddles, kernel spectrum: 64 complex per thread) live in tensor memory;
//     the P2' twiddles are uniform over a half-warp and come from a 4 KB shared table with broadcast reads; the
//     pre/post factors of the window sit in shared memory (32 KB);
//   * the rows of the next pair are loaded into registers right after the last DFT of the current pair, before its
//     post-factor multiply and stores, and the pair after that is prefetched into L2.
#pragma once

#include "cpf_stream_core.h"

namespace cpf {

constexpr int ST_SMEM_BYTES = (2 * ST_GROUP_ELEMS + 256) * (int)sizeof(double2) + 2 * 2048 * (int)sizeof(double);

struct TmemTables {
  uint32_t tb;
  Tm4 buf[2];
  template <int TABLE, int SET>
  __device__ __forceinline__ void issue(const int ch, const int b) { tmem_ld4(tb + 256u * SET + 64u * TABLE + 16u * ch, buf[b]); }
  __device__ __forceinline__ void wait(const int b) { tmem_wait4(buf[b]); }
  template <int SET>
  __device__ __forceinline__ double2 get(const int, const int, const int b, const int i) const { return buf[b].get(i); }
};

template <bool FULLWIN>
__device__ __forceinline__ void st_load_rows(const double* pa, const double* pb, const unsigned m_in, double (&x)[8], double (&y)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (FULLWIN) {
      x[r] = __ldcs(pa + 256 * r);
      y[r] = __ldcs(pb + 256 * r);
    } else {
      const bool ok = (m_in >> r) & 1u;
      x[r] = ok ? __ldcs(pa + 256 * r) : 0.;
      y[r] = ok ? __ldcs(pb + 256 * r) : 0.;
    }
  }
}

// everything the loop needs, precomputed on the host so that it sits in the constant bank instead of registers
struct StreamArgs {
  const double* in;
  double* out;
  const double* pre;        // [P, N]
  const double* post;       // [P, N]
  long long in_row;         // doubles between input rows b and b + 1
  long long in_p;           // doubles between plan rows p and p + 1 of the same b (0 when the input has no P axis)
  long long out_row;        // P * n_out
  long long items;          // P * pairs_per_p
  int n, n_out, P, pairs_per_p;
  int odd_pair;             // index of the pair whose second row does not exist (odd batch), or -1
  int off_in, off_out;      // N/4 - in_left, N/4 - out_left
  int lines;                // 128-byte lines per input row
};

// twtab [256][3][16]: P1, P2, P1' twiddles of thread tau ; uttab [P][256][16]: kernel spectrum at the bins thread tau
// holds after FFT #1 ((-1)^k and 1/N folded in) ; m256 [16][16] = w_256^{h l}
// ABL (lab builds only, -DCPF_LAB): ablation bits — 1: no group barriers (racy), 2: no P2' twiddle loads, 4: no global
// loads/stores, 8: no pre/post-factor loads, 16: no fp64 DFT work is removed (reserved).  Results are wrong on purpose.
template <bool FULLWIN, int ABL = 0>
__global__ void __launch_bounds__(512, 1) fftlog_stream_kernel(const StreamArgs a, const double2* __restrict__ twtab,
                                                               const double2* __restrict__ uttab, const double2* __restrict__ m256) {
  constexpr int T = 256, N = 4096, NG = 2, W = N / 2;
  extern __shared__ double2 smem[];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5;
  const int g = threadIdx.x >> 8, tau = threadIdx.x & 255;
  double2* S = smem + g * ST_GROUP_ELEMS;
  double2* M = smem + NG * ST_GROUP_ELEMS;
  double* spre = reinterpret_cast<double*>(M + 256);
  double* spost = spre + W;

  if (warp == 0) tmem_alloc_all(&s_tmem_base);
  if (threadIdx.x < 256) M[threadIdx.x] = m256[threadIdx.x];
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  TmemTables tb;
  // lane quarter of this warp; threads tau and tau + 128 share a lane and use different column halves; both groups
  // read the same copy
  tb.tb = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u * (uint32_t)(tau >> 7);

  // which of this thread's 8 window elements exist in the unpadded rows
  unsigned m_in = 0xffu, m_out = 0xffu;
  if (!FULLWIN) {
    m_in = m_out = 0u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if ((unsigned)(a.off_in + tau + T * r) < (unsigned)a.n) m_in |= 1u << r;
      if ((unsigned)(a.off_out + tau + T * r) < (unsigned)a.n_out) m_out |= 1u << r;
    }
  }

  const long long lo = a.items * blockIdx.x / gridDim.x, hi = a.items * (blockIdx.x + 1) / gridDim.x;
  bool have_tw = false;

  for (long long seg = lo; seg < hi;) {
    const int p = (int)(seg / a.pairs_per_p);
    const long long seg_hi = min(hi, (long long)(p + 1) * a.pairs_per_p);
    const int pair_lo = (int)(seg - (long long)p * a.pairs_per_p), pair_hi = (int)(seg_hi - (long long)p * a.pairs_per_p);
    seg = seg_hi;
    tmem_fence_before();
    __syncthreads();           // nobody reads the previous plan row's tables any more
    tmem_fence_after();
    if (g == 0) {
      double2 d[4];
      if (!have_tw) {
        const double2* rec = twtab + (size_t)tau * 48;
#pragma unroll 2
        for (int ch = 0; ch < 12; ++ch) {
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = rec[4 * ch + q];
          tmem_st4(tb.tb + 16u * ch + (ch >= 8 ? 64u : 0u), d);     // tables 0, 1 and 3
        }
      }
      const double2* rec = uttab + ((size_t)p * T + tau) * 16;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = rec[4 * ch + q];
        tmem_st4(tb.tb + 64u * ST_UT + 16u * ch, d);
      }
      tmem_wait_st();
    }
    have_tw = true;
    {
      const double* pre = a.pre + (size_t)p * N + N / 4;
      const double* post = a.post + (size_t)p * N + N / 4;
      for (int i = threadIdx.x; i < W; i += 512) { spre[i] = pre[i]; spost[i] = post[i]; }
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    // this thread's first window element of row 2*pair (input) / first output element
    int pair = pair_lo + g;
    const double* pa = a.in + (long long)p * a.in_p + 2LL * pair * a.in_row + (a.off_in + tau);
    double* oa = a.out + 2LL * pair * a.out_row + (long long)p * a.n_out + (a.off_out + tau);
    for (; pair < pair_hi; pair += NG, pa += 2 * NG * a.in_row, oa += 2 * NG * a.out_row) {
      const bool has1 = pair != a.odd_pair;
      {
        double x[8], y[8];
        if (ABL & 4) {
#pragma unroll
          for (int r = 0; r < 8; ++r) { x[r] = 1. + tau; y[r] = 2. + r; }
        } else
        st_load_rows<FULLWIN>(pa, has1 ? pa + a.in_row : pa, m_in, x, y);
        // the pair after next into L2 (one 128-byte line per thread covers both rows)
        const int pf = pair + 2 * NG;
        if (pf < pair_hi && tau < 2 * a.lines) {
          const bool second = tau >= a.lines;
          if (!second || pf != a.odd_pair) {
            const double* q = pa - (a.off_in + tau) + 4 * NG * a.in_row + (second ? a.in_row + 16 * (tau - a.lines) : 16 * tau);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          }
        }
        double2 v8[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { const double pr = (ABL & 8) ? 1.0000001 : spre[tau + T * r]; v8[r] = mk2(x[r] * pr, y[r] * pr); }   // odd tail: y = x, its output is not stored
        st_p1(tau, v8, S, tb);
      }
      if (!(ABL & 1)) named_sync(1 + g, T);
      st_p2(tau, S, tb);
      __syncwarp();
      st_p3_mul_p1(tau, S, tb);
      __syncwarp();
      if (ABL & 2) {
        double2 w[16];
        st_row_load(tau, S, w);
        dft_dit<16, false, false>(w);
        const double2 m = M[16 * (tau >> 4) + 1];
#pragma unroll
        for (int l1 = 1; l1 < 16; ++l1) w[l1] = cmul(w[l1], m);
        st_row_store(tau, w, S);
      } else
      st_p2b(tau, S, M);
      if (!(ABL & 1)) named_sync(1 + g, T);
      double2 v[16];
      st_p3b(tau, v, S);
      double* ob = oa + a.out_row;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (FULLWIN || ((m_out >> r) & 1u)) {
          const double po = (ABL & 8) ? 0.9999999 : spost[tau + T * r];
          if (ABL & 4) { if (v[r].x * po == 1.2345e-300) oa[0] = v[r].y; }
          else {
          __stcs(oa + T * r, v[r].x * po);
          if (has1) __stcs(ob + T * r, v[r].y * po);
          }
        }
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_all(s_tmem_base);
}

// ================================================================================================================
// Two column sets per thread ("stream2"): 128 threads per pair of rows, thread t runs the phases of the virtual
// threads tau = t and tau = t + 128 of cpf_stream_core.h.  NG groups (2 or 3) of 4 warps per CTA: every warp of a
// group sits on a different SM sub-partition, so a sub-partition holds NG warps that belong to NG different pairs of
// rows and never wait for each other.
//
// Why (profiles/r01f_summary.md, r01h ablations): with 256 threads per pair the two warps of a group that share a
// sub-partition leave every barrier together and stay in step, so a sub-partition effectively holds two actors that
// each alternate between fp64 work and shared-memory work, and the fp64 pipe idles whenever both are in a memory
// phase (60 % busy; 83 % with the barriers removed).  Three groups need three exchange buffers (205 KB), which
// leaves no room for the pre/post factors in shared memory: they are read through L1/L2 together with the rows.
// ILV: the two column sets are interleaved by hand (the loads of one set are in flight while the other set's DFT
// runs); otherwise they are processed one after the other.
// ================================================================================================================
template <int NG>
constexpr int st2_smem_bytes() { return (NG * ST_GROUP_ELEMS + 256) * (int)sizeof(double2); }

template <bool FULLWIN>
__device__ __forceinline__ void st2_load_pre(const double* pre, const unsigned m, double (&f)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) f[r] = (FULLWIN || ((m >> r) & 1u)) ? __ldg(pre + 256 * r) : 0.;
}

template <bool FULLWIN, int NG, bool ILV>
__global__ void __launch_bounds__(128 * NG, 1) fftlog_stream2_kernel(const StreamArgs a, const double2* __restrict__ twtab,
                                                                     const double2* __restrict__ uttab, const double2* __restrict__ m256) {
  constexpr int T = 128, N = 4096, NT = T * NG;
  extern __shared__ double2 smem[];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5;
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
  const int tau0 = t, tau1 = t + T;
  double2* S = smem + g * ST_GROUP_ELEMS;
  double2* M = smem + NG * ST_GROUP_ELEMS;

  if (warp == 0) tmem_alloc_all(&s_tmem_base);
  if (threadIdx.x < 256) M[threadIdx.x] = m256[threadIdx.x];
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  TmemTables tb;
  // warps w, w + 4, w + 8 (one per group) share a lane quarter and read the same tables; set 0 in columns [0, 256),
  // set 1 in [256, 512)
  tb.tb = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const double2* M0 = M + 16 * (tau0 >> 4);
  const double2* M1 = M + 16 * (tau1 >> 4);

  unsigned m_in = 0xffffu, m_out = 0xffffu;   // bit 8 s + r: window element tau_s + 256 r exists in the unpadded row
  if (!FULLWIN) {
    m_in = m_out = 0u;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if ((unsigned)(a.off_in + t + T * s + 256 * r) < (unsigned)a.n) m_in |= 1u << (8 * s + r);
        if ((unsigned)(a.off_out + t + T * s + 256 * r) < (unsigned)a.n_out) m_out |= 1u << (8 * s + r);
      }
  }

  const long long lo = a.items * blockIdx.x / gridDim.x, hi = a.items * (blockIdx.x + 1) / gridDim.x;
  bool have_tw = false;

  for (long long seg = lo; seg < hi;) {
    const int p = (int)(seg / a.pairs_per_p);
    const long long seg_hi = min(hi, (long long)(p + 1) * a.pairs_per_p);
    const int pair_lo = (int)(seg - (long long)p * a.pairs_per_p), pair_hi = (int)(seg_hi - (long long)p * a.pairs_per_p);
    seg = seg_hi;
    tmem_fence_before();
    __syncthreads();           // nobody reads the previous plan row's tables any more
    tmem_fence_after();
    if (g == 0) {
      double2 d[4];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int tau = t + T * s;
        if (!have_tw) {
          const double2* rec = twtab + (size_t)tau * 48;
#pragma unroll 2
          for (int ch = 0; ch < 12; ++ch) {
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = rec[4 * ch + q];
            tmem_st4(tb.tb + 256u * s + 16u * ch + (ch >= 8 ? 64u : 0u), d);     // tables 0, 1 and 3
          }
        }
        const double2* rec = uttab + ((size_t)p * 256 + tau) * 16;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = rec[4 * ch + q];
          tmem_st4(tb.tb + 256u * s + 64u * ST_UT + 16u * ch, d);
        }
      }
      tmem_wait_st();
    }
    have_tw = true;
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    int pair = pair_lo + g;
    if (pair >= pair_hi) continue;
    // this thread's first window element of row 2*pair (input) / first output element, and of the pre/post factors
    const double* pa = a.in + (long long)p * a.in_p + 2LL * pair * a.in_row + (a.off_in + t);
    double* oa = a.out + 2LL * pair * a.out_row + (long long)p * a.n_out + (a.off_out + t);
    const double* pre = a.pre + (size_t)p * N + N / 4 + t;
    const double* post = a.post + (size_t)p * N + N / 4 + t;

    for (; pair < pair_hi; pair += NG, pa += 2 * NG * a.in_row, oa += 2 * NG * a.out_row) {
      const bool has1 = pair != a.odd_pair;
      const double* pb = has1 ? pa + a.in_row : pa;
      double2 A[16], B[16];
      // ---- rows -> registers, P1 of both sets (odd tail: y = x, its output is not stored) ----
      {
        double xa[8], ya[8], fa[8], xb[8], yb[8], fb[8];
        st_load_rows<FULLWIN>(pa, pb, m_in, xa, ya);
        st2_load_pre<FULLWIN>(pre, m_in, fa);
        st_load_rows<FULLWIN>(pa + T, pb + T, m_in >> 8, xb, yb);
        st2_load_pre<FULLWIN>(pre + T, m_in >> 8, fb);
        const int pf = pair + 2 * NG;      // the pair after next into L2: two 128-byte lines per thread cover both rows
        if (pf < pair_hi) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int l = t + T * s;
            const bool second = l >= a.lines;
            if (l < 2 * a.lines && (!second || pf != a.odd_pair)) {
              const double* q = pa - (a.off_in + t) + 4 * NG * a.in_row + (second ? a.in_row + 16 * (l - a.lines) : 16 * l);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
          }
        }
        double2 v8[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v8[r] = mk2(xa[r] * fa[r], ya[r] * fa[r]);
        st_p1_compute<TmemTables, 0>(v8, A, tb);
        st_p1_store(tau0, A, S);
#pragma unroll
        for (int r = 0; r < 8; ++r) v8[r] = mk2(xb[r] * fb[r], yb[r] * fb[r]);
        st_p1_compute<TmemTables, 1>(v8, B, tb);
        st_p1_store(tau1, B, S);
      }
      named_sync(1 + g, T);
      // ---- P2, P3, kernel multiply, P1', P2': warp-local ----
      if (ILV) {
        st_row_load(tau0, S, A);
        st_row_load(tau1, S, B);
        st_p2_compute<TmemTables, 0>(A, tb);
        st_row_store(tau0, A, S);
        __syncwarp();
        st_own_load(tau0, S, A);
        st_p2_compute<TmemTables, 1>(B, tb);
        st_row_store(tau1, B, S);
        __syncwarp();
        st_own_load(tau1, S, B);
        {
          double2 A2[16];
          st_p3_compute<TmemTables, 0>(A, A2, tb);
          st_own_store(tau0, A2, S);
        }
        __syncwarp();
        st_row_load(tau0, S, A);
        {
          double2 B2[16];
          st_p3_compute<TmemTables, 1>(B, B2, tb);
          st_own_store(tau1, B2, S);
        }
        __syncwarp();
        st_row_load(tau1, S, B);
        st_p2b_compute(A, M0);
        st_row_store(tau0, A, S);
        st_p2b_compute(B, M1);
        st_row_store(tau1, B, S);
      } else {
        st_row_load(tau0, S, A);
        st_p2_compute<TmemTables, 0>(A, tb);
        st_row_store(tau0, A, S);
        st_row_load(tau1, S, B);
        st_p2_compute<TmemTables, 1>(B, tb);
        st_row_store(tau1, B, S);
        __syncwarp();
        {
          double2 A2[16];
          st_own_load(tau0, S, A);
          st_p3_compute<TmemTables, 0>(A, A2, tb);
          st_own_store(tau0, A2, S);
          st_own_load(tau1, S, B);
          st_p3_compute<TmemTables, 1>(B, A2, tb);
          st_own_store(tau1, A2, S);
        }
        __syncwarp();
        st_row_load(tau0, S, A);
        st_p2b_compute(A, M0);
        st_row_store(tau0, A, S);
        st_row_load(tau1, S, B);
        st_p2b_compute(B, M1);
        st_row_store(tau1, B, S);
      }
      named_sync(1 + g, T);
      // ---- P3', post-factor, store ----
      double fa[8], fb[8];
      st2_load_pre<FULLWIN>(post, m_out, fa);
      st2_load_pre<FULLWIN>(post + T, m_out >> 8, fb);
      st_col_load(tau0, S, A);
      if (ILV) st_col_load(tau1, S, B);
      double* ob = oa + a.out_row;
      st_p3b_compute(A);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (FULLWIN || ((m_out >> r) & 1u)) {
          __stcs(oa + 256 * r, A[r].x * fa[r]);
          if (has1) __stcs(ob + 256 * r, A[r].y * fa[r]);
        }
      }
      if (!ILV) st_col_load(tau1, S, B);
      st_p3b_compute(B);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (FULLWIN || ((m_out >> (8 + r)) & 1u)) {
          __stcs(oa + T + 256 * r, B[r].x * fb[r]);
          if (has1) __stcs(ob + T + 256 * r, B[r].y * fb[r]);
        }
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_all(s_tmem_base);
}

}  // namespace cpf
