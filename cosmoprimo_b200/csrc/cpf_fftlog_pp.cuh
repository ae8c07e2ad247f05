// cpf_fftlog_pp.cuh — persistent "ping-pong" FFTLog kernel (sm_100a) for the default call (zero padding, cropped
// output, real post-factor: cosmoprimo/fftlog.py:198-241 with extrap=0, keep_padding=False).  Included by
// cpf_fftlog.cu; the arithmetic is the register FFT of cpf_fft_core.h, so results are those of fftlog_fast_kernel
// up to the rounding of the twiddle products (the full twiddle set is tabulated here instead of being formed from
// six loaded values).
//
// Why this kernel exists (profiles/r01c_summary.md, tools/lab/fft_lab.cu): in the per-pair kernel the fp64 pipe
// (~4.9 k SM-cycles per pair of rows) and the LSU/shared-memory pipe (~6.5 k: 4.1 k of exchange traffic + table and
// row loads) are both about half busy because the two resident CTAs overlap them only by chance.  Here
//   * one 512-thread CTA per SM stays resident and is split into 512/T groups, each transforming one pair of rows
//     at a time (T = N/16 threads, named barriers per group);
//   * every batch-invariant, thread-private table — the 15+15 inter-pass twiddles, the 16 kernel-spectrum values
//     and the 8+8 pre/post factors a thread needs — lives in TENSOR MEMORY (tcgen05.alloc / .st / .ld, 32x32b
//     shape: one TMEM lane per thread).  TMEM reads do not go through the LSU, run beside LDS at full rate
//     (tools/lab/mio_lab.cu) and cost no fp64 instructions, where the per-pair kernel spends 11 % of its fp64
//     issue slots on rebuilding twiddles and 20 % of its LSU wavefronts on table loads;
//   * the groups run free of each other (handing a "math token" back and forth, or staggering their start, was
//     measured and is not faster: profiles/r01d_pp_driver_first.log);
//   * the rows of the next pair are prefetched into L2 while the current pair is transformed.
#pragma once

#include "cpf_async.h"
#include "cpf_fft_core.h"

namespace cpf {

// ---- tensor-memory helpers (tcgen05, 32x32b shape: thread i of a warp <-> lane 32*(warp%4)+i) ------------------

__device__ __forceinline__ void tmem_alloc_all(uint32_t* slot) {   // whole TMEM (512 columns); one warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pp_smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_all(const uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// four complex doubles = 16 columns of this thread's lane
struct Tm4 {
  uint32_t r[16];
  __device__ __forceinline__ double2 get(const int i) const {
    return mk2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
  }
  __device__ __forceinline__ double getd(const int i) const { return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]); }
};

__device__ __forceinline__ void tmem_ld4(const uint32_t taddr, Tm4& d) {   // asynchronous: tmem_wait4 before use
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(d.r[0]), "=r"(d.r[1]), "=r"(d.r[2]), "=r"(d.r[3]), "=r"(d.r[4]), "=r"(d.r[5]), "=r"(d.r[6]), "=r"(d.r[7]),
        "=r"(d.r[8]), "=r"(d.r[9]), "=r"(d.r[10]), "=r"(d.r[11]), "=r"(d.r[12]), "=r"(d.r[13]), "=r"(d.r[14]), "=r"(d.r[15])
      : "r"(taddr));
}
// the "+r" operands tie the loaded registers to the wait so that no use can be scheduled above it
__device__ __forceinline__ void tmem_wait4(Tm4& d) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(d.r[0]), "+r"(d.r[1]), "+r"(d.r[2]), "+r"(d.r[3]), "+r"(d.r[4]), "+r"(d.r[5]), "+r"(d.r[6]), "+r"(d.r[7]),
                 "+r"(d.r[8]), "+r"(d.r[9]), "+r"(d.r[10]), "+r"(d.r[11]), "+r"(d.r[12]), "+r"(d.r[13]), "+r"(d.r[14]), "+r"(d.r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st4(const uint32_t taddr, const double2 (&d)[4]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[4 * i] = (uint32_t)__double2loint(d[i].x); r[4 * i + 1] = (uint32_t)__double2hiint(d[i].x);
    r[4 * i + 2] = (uint32_t)__double2loint(d[i].y); r[4 * i + 3] = (uint32_t)__double2hiint(d[i].y);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_sync(const int id, const int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier with an OR reduction of `pred` over its threads (barrier.red): returns the OR
__device__ __forceinline__ bool named_sync_or(const int id, const int nthreads, const bool pred) {
  unsigned ret;
  asm volatile(
      "{\n\t.reg .pred pin, pout;\n\tsetp.ne.u32 pin, %3, 0;\n\tbar.red.or.pred pout, %1, %2, pin;\n\tselp.u32 %0, 1, 0, pout;\n\t}"
      : "=r"(ret)
      : "r"(id), "r"(nthreads), "r"((unsigned)pred)
      : "memory");
  return ret != 0;
}
__device__ __forceinline__ void named_arrive(const int id, const int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- per-thread table record (host builds it, see build_pp_tables in cpf_fftlog.cu) ---------------------------
// 64 complex doubles per (plan row p, thread slot t); the first PP_REC_USED are copied to the thread's TMEM lane.
//   [ 0,16)  tw1: pass-1 twiddles  w_N^{n2 k1},  index c*R1 + k1 (n2 = t + T c)
//   [16,32)  tw2: pass-2 twiddles  w_256^{m2 l1}, index l1 (m2 = t % 16)
//   [32,48)  ut : kernel spectrum at the bins t + T r this thread holds after FFT #1 (Hermitian-extended, 1/N and
//                 the (-1)^k of the N/4 rotation folded in)
//   [48,52)  pre [N/4 + t + T r], r = 0..7 (8 doubles)
//   [52,56)  post[N/4 + t + T r], r = 0..7 (8 doubles)
constexpr int PP_REC = 64, PP_REC_USED = 56;
constexpr uint32_t PP_COL_TW1 = 0, PP_COL_TW2 = 64, PP_COL_UT = 128, PP_COL_PRE = 192, PP_COL_POST = 208;

// barrier ids: 0 = __syncthreads, 1..8 = groups
// TMA (full window: n = N/2, in_left = N/4, 16-byte aligned rows): the rows of a group's next pair are staged in shared memory by
// bulk copies issued after the first group barrier of the current pair, as in fftlog_stream_kernel.
// DYN (with TMA): the groups draw their pairs from the plan row's ticket counter (st_draw_ticket in cpf_fftlog_stream.cuh's sense:
// one atomicInc per draw, wrapping to zero after pairs + groups draws) instead of taking every (grid x NG)-th pair.
template <int R1, bool TMA = false, bool DYN = false>
__global__ void __launch_bounds__(512, 1) fftlog_pp_kernel(const FftlogArgs a, const double2* __restrict__ tmtab) {
  static_assert(!DYN || TMA, "dynamic scheduling is built on the staged variant");
  typedef Geo<R1> G;
  constexpr int T = G::T, N = G::N, NG = 512 / T, RS = G::RS;
  extern __shared__ double2 smem[];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(8) uint64_t s_mbar[NG];
  __shared__ long long s_next[NG];
  const int warp = threadIdx.x >> 5;
  const int g = threadIdx.x / T, t = threadIdx.x - g * T;
  double2* S = smem + g * G::SMEM_ELEMS;
  double* stage = reinterpret_cast<double*>(smem + NG * G::SMEM_ELEMS) + g * N;      // TMA only: two rows of N/2 doubles
  unsigned stage_parity = 0;
  if (TMA && t == 0) { mbar_init(&s_mbar[g], 1); mbar_fence_init(); }
  auto stage_rows = [&](const int p, const long long pair) {                        // one thread per group
    const long long b0 = 2 * pair;
    const double* ra = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
    const bool two = b0 + 1 < a.batch;
    mbar_expect_tx(&s_mbar[g], two ? 2u * N * 4u : N * 4u);
    bulk_g2s(stage, ra, N * 4u, &s_mbar[g]);
    if (two) bulk_g2s(stage + N / 2, a.in + (a.in_has_P ? ((b0 + 1) * a.P + p) : b0 + 1) * (long long)a.n, N * 4u, &s_mbar[g]);
  };

  // programmatic dependent launch (see fftlog_stream_kernel): the next kernel of the stream may start its prologue early
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 0) tmem_alloc_all(&s_tmem_base);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  // this thread's lane and column window: warps W, W+4, ... share a lane quarter; for T = 256 the two warp sets of a
  // group (t/32 < 4, >= 4) use different column halves, and both groups read the same copy
  const uint32_t tb = s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (T == 256 ? 256u * (uint32_t)(t >> 7) : 0u);

  auto gbar = [&]() { named_sync(1 + g, T); };

  const long long per_iter = (long long)gridDim.x * NG;
  const long long iters = (a.pairs_per_p + per_iter - 1) / per_iter;

  for (int p = 0; p < a.P; ++p) {
    if (p > 0) { tmem_fence_before(); __syncthreads(); tmem_fence_after(); }   // all reads of the old tables are done
    if (warp < (T / 32 > 4 ? T / 32 : 4)) {
      const double2* rec = tmtab + ((size_t)p * T + t) * PP_REC;
#pragma unroll 2
      for (int ch = 0; ch < PP_REC_USED / 4; ++ch) {
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

    if (p == 0) asm volatile("griddepcontrol.wait;" ::: "memory");      // plan tables only so far; rows may come from the previous kernel
    const unsigned ticket_wrap = (unsigned)(a.pairs_per_p + (long long)gridDim.x * NG - 1);
    long long pair = (long long)blockIdx.x * NG + g;
    if (DYN) {
      if (t == 0) {
        const long long t0 = (long long)atomicInc(a.tickets + p, ticket_wrap);
        s_next[g] = t0;
        if (t0 < a.pairs_per_p) stage_rows(p, t0);
      }
      named_sync(1 + g, T);
      pair = s_next[g];
    } else if (TMA && t == 0 && pair < a.pairs_per_p) stage_rows(p, pair);
    for (long long it = 0; it < (DYN ? (1LL << 62) : iters); ++it) {
      if (!DYN) pair = (it * gridDim.x + blockIdx.x) * NG + g;
      const bool active = pair < a.pairs_per_p;
      if (!active) break;
      const long long b0 = 2 * pair, b1 = b0 + 1;
      const bool has1 = active && b1 < a.batch;
      const double* rowA = a.in + (a.in_has_P ? (b0 * a.P + p) : b0) * (long long)a.n;
      const double* rowB = has1 ? a.in + (a.in_has_P ? (b1 * a.P + p) : b1) * (long long)a.n : rowA;

      // L2 prefetch of the rows this group transforms next
      if (!TMA) {
        const long long nb0 = 2 * (pair + per_iter);
        if (nb0 < a.batch) {
          const int lines = (a.n * 8 + 127) / 128;
          const double* nA = a.in + (a.in_has_P ? (nb0 * a.P + p) : nb0) * (long long)a.n;
          const double* nB = nb0 + 1 < a.batch ? a.in + (a.in_has_P ? ((nb0 + 1) * a.P + p) : nb0 + 1) * (long long)a.n : nA;
          for (int l = t; l < 2 * lines; l += T) {
            const double* q = (l < lines ? nA : nB) + (size_t)(l < lines ? l : l - lines) * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          }
        }
      }

      double2 v[16];
      Tm4 tq;
      bool bad_a = false, bad_b = false;     // non-finite samples are zeroed; their row is written as NaN (see scrub)
      // ---- load the two rows (window [N/4, 3N/4) of the padded row), times pre ----
      {
        tmem_ld4(tb + PP_COL_PRE, tq);
        double x[8], y[8];
        if (TMA) {
          mbar_wait(&s_mbar[g], stage_parity);
          stage_parity ^= 1u;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            x[r] = stage[t + T * r];
            y[r] = has1 ? stage[N / 2 + t + T * r] : 0.;
          }
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int i = t + T * r + N / 4 - a.in_left;
            const bool ok = active && (unsigned)i < (unsigned)a.n;
            x[r] = ok ? __ldcs(rowA + i) : 0.;
            y[r] = (ok && has1) ? __ldcs(rowB + i) : 0.;
          }
        }
        scrub_rows(x, y, bad_a, bad_b);
        tmem_wait4(tq);
#pragma unroll
        for (int r = 0; r < 8; ++r) { const double pr = tq.getd(r); v[r] = mk2(x[r] * pr, y[r] * pr); }
#pragma unroll
        for (int r = 8; r < 16; ++r) v[r] = mk2(0., 0.);
      }

      // a register pass followed by its twiddles (chunks of 4 from TMEM, next chunk in flight while this one is used)
      // and the scatter to shared memory
      auto twiddle_store = [&](double2 (&w)[16], const int nw, const uint32_t col, double2* dst, const int stride, auto dft) {
        // dft() turns w[0..nw) into the DFT outputs (the first chunk of twiddles arrives behind its butterflies); w[k] *= tw[k] for k >= 1;
        // then dst[k * stride] = w[k]
        Tm4 tw[2];
        tmem_ld4(col, tw[0]);
        dft();
        tmem_wait4(tw[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (4 * ch < nw) {
            if (4 * (ch + 1) < nw) tmem_ld4(col + 16 * (ch + 1), tw[(ch + 1) & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = 4 * ch + i;
              if (k > 0 && k < nw) w[k] = cmul(w[k], tw[ch & 1].get(i));
            }
            if (4 * (ch + 1) < nw) tmem_wait4(tw[(ch + 1) & 1]);
          }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (k < nw) dst[k * stride] = w[k];
      };

      auto pass1 = [&](auto half_in) {
        constexpr bool HALF_IN = decltype(half_in)::value;
#pragma unroll
        for (int c = 0; c < G::C; ++c) {
          const int n2 = t + T * c;
          double2 w[16];
#pragma unroll
          for (int n1 = 0; n1 < R1; ++n1) w[bitrev(n1, G::B1)] = v[n1 * G::C + c];
          twiddle_store(w, R1, tb + PP_COL_TW1 + 4 * (c * R1), S + n2, RS, [&]() { dft_dit<R1, HALF_IN, false>(w); });
        }
      };
      auto pass2 = [&]() {
        const int k1 = t >> 4, m2 = t & 15;
        double2* row = S + k1 * RS + m2;
        double2 w[16];
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) w[bitrev(m1, 4)] = row[16 * m1];
        twiddle_store(w, 16, tb + PP_COL_TW2, row, 16, [&]() { dft_dit<16, false, false>(w); });
      };

      // ---- FFT #1 ----
      pass1(std::true_type());
      const bool row_a_bad = named_sync_or(1 + g, T, bad_a);
      if (TMA && t == 0) {                                                                   // the staged rows are in registers everywhere
        if (DYN) {
          const long long tn = (long long)atomicInc(a.tickets + p, ticket_wrap);
          s_next[g] = tn;                                                                      // read by the group after its next barrier
          if (tn < a.pairs_per_p) stage_rows(p, tn);
        } else if (pair + per_iter < a.pairs_per_p) stage_rows(p, pair + per_iter);
      }
      pass2();
      const bool row_b_bad = named_sync_or(1 + g, T, bad_b);
      const long long next_pair = DYN ? s_next[g] : 0;
      {
        const int k1 = t & (R1 - 1), l1 = t >> G::B1;
        const double2* row = S + k1 * RS + 16 * l1;
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = row[m2];
      }
      gbar();   // pass-3 reads of S are done before FFT #2 overwrites it (nothing below touches S before pass 1 stores)
      // ---- kernel multiply ----
      {
        Tm4 tu[2];
        tmem_ld4(tb + PP_COL_UT, tu[0]);
        dft_dit<16, false, false>(v);
        tmem_wait4(tu[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (ch < 3) tmem_ld4(tb + PP_COL_UT + 16 * (ch + 1), tu[(ch + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[4 * ch + i] = cmul(v[4 * ch + i], tu[ch & 1].get(i));
          if (ch < 3) tmem_wait4(tu[(ch + 1) & 1]);
        }
      }
      // ---- FFT #2 ----
      pass1(std::false_type());
      gbar();
      pass2();
      gbar();
      {
        const int k1 = t & (R1 - 1), l1 = t >> G::B1;
        const double2* row = S + k1 * RS + 16 * l1;
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) v[bitrev(m2, 4)] = row[m2];
      }
      gbar();   // S is free for the next pair
      tmem_ld4(tb + PP_COL_POST, tq);
      dft_dit<16, false, true>(v);
      tmem_wait4(tq);
#pragma unroll
      for (int r = 0; r < 8; ++r) { const double po = tq.getd(r); v[r].x *= po; v[r].y *= po; }
      if (active) {
        double* outA = a.out + (size_t)(b0 * a.P + p) * a.n_out;
        double* outB = a.out + (size_t)(b1 * a.P + p) * a.n_out;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int o = t + T * r + N / 4 - a.out_left;
          if ((unsigned)o < (unsigned)a.n_out) {
            __stcs(outA + o, row_a_bad ? nan("") : v[r].x);
            if (has1) __stcs(outB + o, row_b_bad ? nan("") : v[r].y);
          }
        }
      }
      if (DYN) pair = next_pair;
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (DYN) ticket_release(a.t_finished, a.t_done, a.t_seq);
  if (warp == 0) tmem_dealloc_all(s_tmem_base);
}

}  // namespace cpf
