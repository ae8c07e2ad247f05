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
//     threads tau and tau + 128 share a lane and the same P2 twiddles, which makes room for the 8 + 8 pre/post factors
//     of a thread in tensor memory too; the P2' twiddles are uniform over a half-warp and come from a 4 KB shared
//     table with broadcast reads;
//   * full window (n = N/2): the two rows of a group's next pair arrive in a shared-memory staging buffer through TMA bulk
//     copies issued one pair ahead (template parameter TMA); other windows load their samples directly, with the pair after
//     next prefetched into L2;
//   * the groups of the CTAs that share a plan row draw their pairs from a self-resetting ticket counter (DYN), which evens
//     out the different speeds of the SMs; small launches keep the static split;
//   * launched with programmatic stream serialisation: the prologue (TMEM allocation, tables) overlaps the tail of the previous
//     kernel of the stream, `griddepcontrol.wait` precedes the first access to caller data;
//   * non-finite samples are zeroed on load and their row is written as NaN (two rows share one complex FFT).
#pragma once

#include "cpf_stream_core.h"

namespace cpf {

constexpr int ST_SMEM_BYTES = (2 * ST_GROUP_ELEMS + 256) * (int)sizeof(double2);
// TMA variant (full window, n = 2048): + one staging buffer of two input rows per group, filled by bulk copies one pair ahead
constexpr int ST_STAGE_DOUBLES = 2 * 2048;
constexpr int ST_SMEM_BYTES_TMA = ST_SMEM_BYTES + 2 * ST_STAGE_DOUBLES * (int)sizeof(double);

// Tensor-memory layout of one lane (512 columns of 32 bits = 256 doubles).  Threads tau and tau + 128 of a group share a lane (and have
// the same L = tau % 16, hence the same P2 twiddles and P3 ratios); both groups read the same copy.  Regions as in cpf_stream_core.h:
//   [  0,  64)  ST_TW2 : t of the P2 twiddles w_256^{L l1} + ratios of the P3 DFT          (shared by the two halves)
//   half h = tau / 128 at 64 + 224 h:
//   [+  0,+ 64) ST_TW1 : t of the P1 twiddles w_4096^{tau k1} + ratios of the P2 DFT
//   [+ 64,+128) ST_UT  : kernel spectrum at the bins H + 16 L + 256 l2
//   [+128,+192) ST_TW1B: t of the P1' twiddles w_4096^{(H + 16 L) k1'} + ratios of the P2' DFT
//   [+192,+208) pre-factor  at window elements tau + 256 r, r < 8   (8 doubles)
//   [+208,+224) post-factor at window elements tau + 256 r, r < 8   (8 doubles)
constexpr uint32_t ST_COL_HALF0 = 64, ST_COL_HALF = 224, ST_COL_PRE = 192, ST_COL_POST = 208;
__host__ __device__ constexpr uint32_t st_table_col(const int table) {
  return table == ST_TW1 ? 0u : table == ST_UT ? 64u : 128u;     // offset inside the half; ST_TW2 is the shared block
}

struct TmemTables {
  uint32_t lane;   // lane quarter of this warp, column 0
  uint32_t half;   // this thread's half
  Tm4 buf[2];
  template <int TABLE, int SET>
  __device__ __forceinline__ void issue(const int ch, const int b) {
    tmem_ld4((TABLE == ST_TW2 ? lane : half + st_table_col(TABLE)) + 16u * ch, buf[b]);
  }
  __device__ __forceinline__ void wait(const int b) { tmem_wait4(buf[b]); }
  template <int SET>
  __device__ __forceinline__ double getd(const int, const int, const int b, const int i) const { return buf[b].getd(i); }
  template <int SET>
  __device__ __forceinline__ double2 get(const int, const int, const int b, const int i) const { return buf[b].get(i); }
};

template <bool FULLWIN>
__device__ __forceinline__ void st_load_rows(const double* pa, const double* pb, const unsigned m_in, double (&x)[8], double (&y)[8],
                                             bool& bad_a, bool& bad_b) {
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
  // non-finite samples are zeroed here and their row is written as NaN at the end (see scrub in cpf_fftlog.cu)
  scrub_rows(x, y, bad_a, bad_b);
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
  long long* dbg;           // lab builds: per-CTA time stamps (globaltimer ns), else null
  int wskew_inv;            // ... the first-barrier offset applies to the OTHER warps (default)
  int wskew_mask;           // phase offset inside a group (launch_stream): warps with (warp index & mask) != 0 idle ...
  int wskew2_ns;            // ... this long after the second group barrier of a pair (default 250 ns)
  int wskew_ns;             // ... and this long after the first (default 300 ns, on the other warps)
  int skew_ns;              // lab builds: group 1 starts its first pair this much later (phase offset between the groups)
  int early;                // the first ticket is drawn and the first pair's rows are prefetched into L2 BEFORE the dependency wait (default 1)
  unsigned* tickets;        // dynamic scheduling (TMA variant, grid >= P): one self-resetting ticket counter per plan row, else null
  unsigned* t_finished;     // ... with the slot's CTA exit counter and completion word (ticket_release, cpf_fftlog.cu)
  unsigned* t_done;
  unsigned t_seq;
};

// Work split: when there are at least as many CTAs as plan rows, every CTA works on ONE plan row (its tables are
// loaded once) and the CTAs of a plan row share its pairs evenly; otherwise CTAs take contiguous item ranges.
// (Drawing pairs from a per-plan-row atomic counter instead was measured and is not faster: r01n.)
__device__ __forceinline__ void st_item_range(const StreamArgs& a, long long& lo, long long& hi, int& row_ctas) {
  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  row_ctas = 0;
  if (G >= a.P) {
    const int base = G / a.P, rem = G % a.P;     // the first `rem` plan rows get base + 1 CTAs
    int p, j, c;
    if (b < rem * (base + 1)) { p = b / (base + 1); j = b - p * (base + 1); c = base + 1; }
    else { const int bb = b - rem * (base + 1); p = rem + bb / base; j = bb - (p - rem) * base; c = base; }
    row_ctas = c;
    lo = (long long)p * a.pairs_per_p + (long long)a.pairs_per_p * j / c;
    hi = (long long)p * a.pairs_per_p + (long long)a.pairs_per_p * (j + 1) / c;
  } else {
    lo = a.items * b / G;
    hi = a.items * (b + 1) / G;
  }
}

#ifdef CPF_LAB
#define ST_STAMP(slot)                                                                              \
  do {                                                                                              \
    if (a.dbg && threadIdx.x == 0) {                                                                \
      long long t_;                                                                                 \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
      a.dbg[8 * blockIdx.x + (slot)] = t_;                                                          \
    }                                                                                               \
  } while (0)
#else
#define ST_STAMP(slot) do { } while (0)
#endif

// Next pair of a plan row's queue.  One hardware atomic per draw (a compare-and-swap loop collapses under the contention
// of ~300 groups: measured 5x slower overall).  The counter wraps to zero by itself: the groups that share the row make
// exactly pairs + groups draws in a launch (every group draws until its first ticket >= pairs), and atomicInc(word, wrap)
// returns to 0 after wrap + 1 draws, so no launch has to reset it.
__device__ __forceinline__ int st_draw_ticket(unsigned* word, const unsigned wrap) { return (int)atomicInc(word, wrap); }

// twtab [3][32][256] doubles: the ST_TW1, ST_TW2, ST_TW1B regions of cpf_stream_core.h (st_build_tables; entry-major: thread tau reads
// element [.][.][tau], coalesced) ;
// uttab [P][16][256]: kernel spectrum at the bins thread tau holds after FFT #1 ((-1)^k and 1/N folded in) ;
// m256 [2][16][16] doubles: t of the P2' twiddles w_256^{h l} and the ratios of the P3' DFT (uniform over a half-warp: shared memory)
// ABL (lab builds only, -DCPF_LAB): ablation bits — 1: no group barriers (racy), 4: no global loads/stores, 8: no pre/post-factor
// loads (results are wrong on purpose); 16: L1 prefetch of the group's next pair, 32: L2 prefetch three pairs ahead instead of two
// (results unchanged).
// TMA (full window only): the two rows of a group's next pair are brought into a shared-memory staging buffer by bulk copies
// issued by one thread right after the first group barrier of the current pair (every thread has consumed the staged rows by
// then), a whole pair ahead of their use; threads then read their 8 + 8 samples with conflict-free 64-bit shared loads instead
// of waiting for global loads.
template <bool FULLWIN, int ABL = 0, bool TMA = false, bool DYN = false>
__global__ void __launch_bounds__(512, 1) fftlog_stream_kernel(const StreamArgs a, const double* __restrict__ twtab,
                                                               const double2* __restrict__ uttab, const double2* __restrict__ m256) {
  static_assert(!TMA || FULLWIN, "the staged variant covers the full window only");
  static_assert(!DYN || TMA, "dynamic scheduling is built on the staged variant");
  constexpr int T = 256, N = 4096, NG = 2;
  extern __shared__ double2 smem[];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(8) uint64_t s_mbar[NG];
  const int warp = threadIdx.x >> 5;
  const int g = threadIdx.x >> 8, tau = threadIdx.x & 255;
  double2* S = smem + g * ST_GROUP_ELEMS;
  double2* M = smem + NG * ST_GROUP_ELEMS;
  const double* Md = reinterpret_cast<const double*>(M);
  double* stage = reinterpret_cast<double*>(M + 256) + g * ST_STAGE_DOUBLES;     // TMA only
  unsigned stage_parity = 0;
  if (TMA && tau == 0) { mbar_init(&s_mbar[g], 1); mbar_fence_init(); }
  // bulk copies of the rows of pair `pair` of plan row p into this group's staging buffer (one thread)
  auto stage_rows = [&](const int p, const int pair) {
    const double* ra = a.in + (long long)p * a.in_p + 2LL * pair * a.in_row;
    const bool two = pair != a.odd_pair;
    mbar_expect_tx(&s_mbar[g], two ? 2u * N * 4u : N * 4u);          // a row is N/2 doubles = 4 N bytes
    bulk_g2s(stage, ra, N * 4u, &s_mbar[g]);
    if (two) bulk_g2s(stage + N / 2, ra + a.in_row, N * 4u, &s_mbar[g]);
  };

  __shared__ int s_next[NG];
  // dynamic scheduling: the groups of the CTAs that share a plan row draw its pairs from a queue, one pair ahead of their use
  // (the draw decides what the bulk copies fetch); evens out the 10 % spread of the CTA end times of the static split
  constexpr bool dynamic = DYN;
  ST_STAMP(0);
  // programmatic dependent launch: let the next kernel of the stream start its CTAs as SMs become free (its prologue -- TMEM
  // allocation, plan tables -- then overlaps the tail of this grid); it waits below, before touching caller data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  long long lo, hi;
  int row_ctas;
  st_item_range(a, lo, hi, row_ctas);
  const unsigned ticket_wrap = (unsigned)(a.pairs_per_p + NG * row_ctas - 1);
  if (dynamic) {        // this CTA's plan row (the host guarantees grid >= P and non-empty static ranges): the whole row is one segment
    const int prow = (int)(lo / a.pairs_per_p);
    lo = (long long)prow * a.pairs_per_p;
    hi = lo + a.pairs_per_p;
  }

  // L2 prefetch of the two rows of pair `pair` of plan row p (one 128-byte line per thread covers both rows)
  auto prefetch_rows = [&](const int p, const int pair, const bool l1 = false) {
    if (tau < 2 * a.lines) {
      const bool second = tau >= a.lines;
      if (!second || pair != a.odd_pair) {
        const double* q = a.in + (long long)p * a.in_p + (2LL * pair + (second ? 1 : 0)) * a.in_row + 16 * (second ? tau - a.lines : tau);
        if (l1) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
      }
    }
  };

  // batch-invariant twiddles: group 0 brings P1 and P2, group 1 brings P1' (the same lanes are visible to both);
  // the loads are in flight while tensor memory is being allocated
  double2 d[4][4];
  const double* rec = twtab + (g == 0 ? 0 : 2 * 32 * T) + tau;          // group 0: region 0 (then 1), group 1: region 2
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
#pragma unroll
    for (int q = 0; q < 4; ++q) d[ch][q] = mk2(rec[(8 * ch + 2 * q) * T], rec[(8 * ch + 2 * q + 1) * T]);
  if (warp == 0) tmem_alloc_all(&s_tmem_base);
  if (threadIdx.x < 256) M[threadIdx.x] = m256[threadIdx.x];
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  TmemTables tb;
  tb.lane = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  tb.half = tb.lane + ST_COL_HALF0 + ST_COL_HALF * (uint32_t)(tau >> 7);
  {
    const uint32_t col = tb.half + st_table_col(g == 1 ? ST_TW1B : ST_TW1);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) tmem_st4(col + 16u * ch, d[ch]);
    if (g == 0) {    // the shared P2 block (both halves write the same values)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int q = 0; q < 4; ++q) d[ch][q] = mk2(rec[(32 + 8 * ch + 2 * q) * T], rec[(32 + 8 * ch + 2 * q + 1) * T]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_st4(tb.lane + 16u * ch, d[ch]);
    }
  }
  ST_STAMP(1);

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
  bool first_seg = true, first_pair = true;
#ifdef CPF_LAB
  if (a.skew_ns > 0 && g == 1) __nanosleep((unsigned)a.skew_ns);
#endif

  for (long long seg = lo; seg < hi;) {
    const int p = (int)(seg / a.pairs_per_p);
    const long long seg_hi = min(hi, (long long)(p + 1) * a.pairs_per_p);
    const int pair_lo = (int)(seg - (long long)p * a.pairs_per_p), pair_hi = (int)(seg_hi - (long long)p * a.pairs_per_p);
    seg = seg_hi;
    int pair = pair_lo + g;
    if (!TMA && pair < pair_hi) prefetch_rows(p, pair);
    if (!first_seg) {
      tmem_fence_before();
      __syncthreads();           // nobody reads the previous plan row's tables any more
      tmem_fence_after();
    }
    first_seg = false;
    if (g == 0) {              // plan-row tables: kernel spectrum, pre- and post-factor
      const double2* ut = uttab + (size_t)p * 16 * T + tau;
      const double* pre = a.pre + (size_t)p * N + N / 4 + tau;
      const double* post = a.post + (size_t)p * N + N / 4 + tau;
      double2 e[2][4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int q = 0; q < 4; ++q) d[ch][q] = ut[(4 * ch + q) * T];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        e[0][q] = mk2(pre[T * (2 * q)], pre[T * (2 * q + 1)]);
        e[1][q] = mk2(post[T * (2 * q)], post[T * (2 * q + 1)]);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_st4(tb.half + st_table_col(ST_UT) + 16u * ch, d[ch]);
      tmem_st4(tb.half + ST_COL_PRE, e[0]);
      tmem_st4(tb.half + ST_COL_POST, e[1]);
    }
    tmem_wait_st();
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    ST_STAMP(2);
    // everything above read plan tables only; rows may have been written by the previous kernel of the stream.  What does not READ caller
    // data happens before the dependency wait, while the previous grid drains: the first ticket (the counters belong to this launch) and an L2
    // prefetch of the first pair's rows (L2 is the coherence point: a line prefetched early still shows every later write of the previous
    // grid); the bulk copies themselves follow the wait.
    if (first_pair && !a.early) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (dynamic) {
      if (tau == 0) s_next[g] = st_draw_ticket(a.tickets + p, ticket_wrap);
      named_sync(1 + g, T);
      pair = s_next[g];
    }
    if (first_pair && a.early) {
      if (TMA && pair < pair_hi) prefetch_rows(p, pair);
      asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (TMA && tau == 0 && pair < pair_hi) stage_rows(p, pair);     // nobody reads the staging buffer any more (barrier above)

    while (pair < pair_hi) {
      const bool has1 = pair != a.odd_pair;
      // this thread's first window element of row 2*pair (input) / first output element
      const double* pa = a.in + (long long)p * a.in_p + 2LL * pair * a.in_row + (a.off_in + tau);
      double* oa = a.out + 2LL * pair * a.out_row + (long long)p * a.n_out + (a.off_out + tau);
      bool bad_a = false, bad_b = false, row_a_bad = false, row_b_bad = false;
      {
        double x[8], y[8];
        if (ABL & 4) {
#pragma unroll
          for (int r = 0; r < 8; ++r) { x[r] = 1. + tau; y[r] = 2. + r; }
        } else if (TMA) {
          mbar_wait(&s_mbar[g], stage_parity);
          stage_parity ^= 1u;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            x[r] = stage[tau + 256 * r];
            y[r] = has1 ? stage[N / 2 + tau + 256 * r] : x[r];
          }
          scrub_rows(x, y, bad_a, bad_b);
        } else
        st_load_rows<FULLWIN>(pa, has1 ? pa + a.in_row : pa, m_in, x, y, bad_a, bad_b);
        Tm4 tf;
        tmem_ld4(tb.half + ST_COL_PRE, tf);
        if (!TMA && pair + ((ABL & 32) ? 3 : 2) * NG < pair_hi) prefetch_rows(p, pair + ((ABL & 32) ? 3 : 2) * NG);
        if ((ABL & 16) && pair + NG < pair_hi) prefetch_rows(p, pair + NG, true);      // lab: the group's next pair into L1
        tmem_wait4(tf);
        double2 v8[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { const double pr = (ABL & 8) ? 1.0000001 : tf.getd(r); v8[r] = mk2(x[r] * pr, y[r] * pr); }   // odd tail: y = x, its output is not stored
        st_p1(tau, v8, S, tb);
      }
      if (!(ABL & 1)) row_a_bad = named_sync_or(1 + g, T, bad_a);
      if (a.wskew_ns > 0 && ((warp & a.wskew_mask) != 0) != (a.wskew_inv != 0)) __nanosleep((unsigned)a.wskew_ns);
      if (TMA && tau == 0) {                                                  // every thread of the group has its samples in registers
        if (dynamic) {
          const int tn = st_draw_ticket(a.tickets + p, ticket_wrap);
          s_next[g] = tn;                                                       // read by the group after its second barrier
          if (tn < pair_hi) stage_rows(p, tn);
        } else if (pair + NG < pair_hi) stage_rows(p, pair + NG);
      }
      st_p2(tau, S, tb);
      __syncwarp();
      st_p3_mul_p1(tau, S, tb);
      __syncwarp();
      st_p2b(tau, S, tb, Md);
      if (!(ABL & 1)) row_b_bad = named_sync_or(1 + g, T, bad_b);
      if (a.wskew2_ns > 0 && (warp & a.wskew_mask)) __nanosleep((unsigned)a.wskew2_ns);   // takes the two warps of a sub-partition out of step
      const int next_pair = dynamic ? s_next[g] : pair + NG;
      double2 v[16];
      st_col_load(tau, S, v);
      Tm4 tf;
      tmem_ld4(tb.half + ST_COL_POST, tf);
      st_p3b_compute(v, Md + 256 + 16 * (tau >> 4));
      tmem_wait4(tf);
      double* ob = oa + a.out_row;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (FULLWIN || ((m_out >> r) & 1u)) {
          const double po = (ABL & 8) ? 0.9999999 : tf.getd(r);
          if (ABL & 4) { if (v[r].x * po == 1.2345e-300) oa[0] = v[r].y; }
          else {
            __stcs(oa + T * r, row_a_bad ? nan("") : v[r].x * po);
            if (has1) __stcs(ob + T * r, row_b_bad ? nan("") : v[r].y * po);
          }
        }
      }
      if (first_pair) { ST_STAMP(3); first_pair = false; }
      pair = next_pair;
    }
    ST_STAMP(4);
  }
  ST_STAMP(5);
  tmem_fence_before();
  __syncthreads();
  if (dynamic) ticket_release(a.t_finished, a.t_done, a.t_seq);      // every draw of this CTA is done
  if (warp == 0) tmem_dealloc_all(s_tmem_base);
  ST_STAMP(6);
}

}  // namespace cpf
