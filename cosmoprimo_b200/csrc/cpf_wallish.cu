// cpf_wallish.cu — Wallish2018 no-wiggle filter (cosmoprimo/bao_filter.py:361-431) and the orthonormal DST-II/III it is
// built on (scipy.fftpack.dst / idst, bao_filter.py:372, 412), on sm_100a.
//
// Pipeline of cpf_wallish2018 for ncols spectra (reference layout: wavenumber along axis 0, one column per spectrum):
//   1. wallish_fused_kernel  — one CTA per PAIR of columns, everything in shared memory (cpf_wallish_core.h):
//        log(k P) -> DST-II (one packed complex FFT-4096, the FFTLog register FFT) -> clamped-spline second derivatives
//        of the even / odd coefficients -> argmax boxes -> cut + re-spline -> DST-III -> exp(.)/k on 1e-2 < k < 1.5,
//        written straight into the knot-value matrix of the final spline.
//   2. wallish_edges_kernel  — rows of that matrix taken from the unfiltered spectrum (k < 5e-4 and k > 2, :415-419).
//   3. spline_factor_kernel + spline_solve_kernel (cpf_spline.cu) — clamped spline on the 3666 shared knots (:420).
//   4. wallish_final_kernel  — evaluate at self.k, blend with the Gaussian top-hat (:421-423).
#include <math.h>
#include <mutex>
#include <vector>

#include "cpf_common.h"
#include "cpf_fft_core.h"
#include "cpf_spline_core.h"
#include "cpf_wallish_core.h"

namespace cpf {

// defined in cpf_spline.cu
int spline_fit_device(const double* d_x, const double* d_y, int nx, long long ncols, int bc, double* d_s, double* d_fac,
                      cudaStream_t stream, bool fac_ready);
void spline_factor_host(const double* x, int nx, int bc, double* fac);

struct WallishTables {
  int device;
  double2* tw1;    // [6,256] factored FFT twiddles, N = 4096
  double2* tw2;    // [6,16]
  double2* twd;    // [4096] exp(-i pi k / 2N)
  double* wtab;    // [32] Thomas pivots of the clamped uniform system
};

struct WallishArgs {
  const double* klin;     // [4096]
  const double* pklin;    // [4096, ncols]
  double2* packed;        // [npairs, 4096]: in  = sign * log(k P) of both columns of a pair in Makhoul order (wallish_pack_kernel),
                          //                 out = raw FFT bins of the DST-III (consumed by wallish_unpack_kernel)
  long long ncols;        // columns of this chunk
  long long ld;           // doubles between consecutive rows of pklin / pkout / pknow (= total number of columns)
  int i0, i1;             // rows of klin kept (1e-2 < k < 1.5)
  int nl;                 // rows of the knot matrix before them
  double* vals;           // [nknots, ncols] knot values of the final spline
  int* boxes;             // [ncols, 4] or null
  const double2 *tw1, *tw2, *twd;
  const double* wtab;
};

// shared-memory carve-up of the fused kernels
struct WallishSmem {
  double2* S;     // FFT exchange buffer, then natural-order spectrum, then reduced right-hand sides
  double2* X;     // DST coefficients, de-interleaved + padded
  double2* DD;    // second derivatives
  double* red;    // [256]
  int* redi;      // [256]
  int* box;       // [8]
  WallishGap* gaps;   // [4]
  double* wtab;   // [32] Thomas pivots (copied from global memory once: they sit in the dependency chain of the first chunks)
  double2* E;     // [256] chunk results of the two-step eliminations (value with zero inflow)
  double* Mf;     // [256] ... and the factor the inflow is multiplied with
  __device__ explicit WallishSmem(double2* base) {
    S = base; X = base + WallishGeo::BUF; DD = base + 2 * WallishGeo::BUF;
    red = reinterpret_cast<double*>(base + 3 * WallishGeo::BUF);
    redi = reinterpret_cast<int*>(red + 256);
    box = redi + 256;
    gaps = reinterpret_cast<WallishGap*>(box + 8);
    wtab = reinterpret_cast<double*>(gaps + 4);
    E = reinterpret_cast<double2*>(wtab + 32);
    Mf = reinterpret_cast<double*>(E + 256);
  }
};
static constexpr size_t kWallishSmemBytes = 3 * (size_t)WallishGeo::BUF * sizeof(double2) + 256 * sizeof(double) + 264 * sizeof(int) + 4 * sizeof(WallishGap) + 32 * sizeof(double) + 256 * sizeof(double2) + 256 * sizeof(double);

__device__ __forceinline__ void fft4096(const int t, double2 (&v)[16], double2* S, const double2* tw1, const double2* tw2) {
  fft_pass1<16, false>(t, v, S, tw1);
  __syncthreads();
  fft_pass2<16>(t, S, tw2);
  __syncthreads();
  fft_pass3<16, false>(t, v, S);
}

// DST-II (orthonormal) of the two packed real sequences whose Makhoul-permuted samples are in v; result in sm.X
__device__ __forceinline__ void dst2_in_smem(const int t, double2 (&v)[16], const WallishSmem& sm, const double2* tw1,
                                             const double2* tw2, const double2* twd) {
  fft4096(t, v, sm.S, tw1, tw2);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) sm.S[t + 256 * r] = v[r];
  __syncthreads();
  wallish_dst2_post(t, v, sm.S, sm.X, twd);
  __syncthreads();
}

// Makhoul input position: FFT sample n takes x'[j], x' = (-1)^j x
__device__ __forceinline__ int makhoul_src(const int n, double& sign) {
  if (n < WallishGeo::N / 2) { sign = 1.; return 2 * n; }
  sign = -1.;
  return 2 * (WallishGeo::N - 1 - n) + 1;
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

__global__ void __launch_bounds__(256, 1) wallish_fused_kernel(const WallishArgs a) {
  typedef WallishGeo G;
  extern __shared__ double2 smem_raw[];
  const WallishSmem sm(smem_raw);
  const int t = threadIdx.x;
  const long long col0 = 2LL * blockIdx.x;
  const bool has1 = col0 + 1 < a.ncols;
  double2 v[16];
  // sign * log(k P) in Makhoul order, packed per pair by wallish_pack_kernel                  (bao_filter.py:371)
  double2* zrow = a.packed + (long long)blockIdx.x * G::N;
  if (t < 32) sm.wtab[t] = a.wtab[t];            // visible after the barriers of the first FFT
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = __ldcs(zrow + t + 256 * r);
  dst2_in_smem(t, v, sm, a.tw1, a.tw2, a.twd);                                              // :372
  // second derivatives of the clamped splines through the even / odd coefficients           (:377-382)
  wallish_forward_local(t, sm.X, sm.E, sm.Mf, sm.wtab);
  __syncthreads();
  wallish_forward_store(t, sm.X, sm.S, sm.E, sm.Mf, sm.wtab);
  __syncthreads();
  wallish_backward_local(t, sm.S, sm.E, sm.Mf, sm.wtab);
  __syncthreads();
  WallishBest chunk;
  wallish_backward_dd(t, sm.X, sm.S, sm.DD, sm.E, sm.Mf, sm.wtab, &chunk);
  // boxes (:392-395): argmax over [20, H-20), then over [first + 5, H-20); per-chunk maxima come out of the backward pass,
  // warps reduce them with shuffles, thread q < 4 merges the four warps of its sequence
  wallish_best_reduce(t, chunk, sm.red, sm.redi);
  __syncthreads();
  if (t < 4) sm.box[2 * t] = wallish_best_final(t, sm.red, sm.redi);
  __syncthreads();
  {
    const int h = t >> 7;
    const WallishBest cand = wallish_chunk_candidate(t, sm.DD, sm.box[4 * h] + G::MARGIN_SECOND, sm.box[4 * h + 2] + G::MARGIN_SECOND, chunk);
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
  // composed in row order by a shuffle tree (5 rounds) instead of a 32-step chain on one thread (14 % of the kernel's stall
  // samples sat on the barrier behind those chains).  Then the 2x2 solves on 4 threads.
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
        fb = wallish_gap_rhs(sm.X, q >> 1, q & 1, i) * w;
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
  if (t < 4) sm.gaps[t] = wallish_gap_finish(sm.X, t >> 1, t & 1, sm.box[2 * t], sm.box[2 * t + 1], sm.red[2 * t], sm.red[2 * t + 1], sm.wtab);
  __syncthreads();
  {                                                                                         // :402
    // only the knots of the removed box change (64 threads per sequence); a box that reaches the end of the array
    // leaves NaN from b0 on, as the reference's spline does beyond its last knot
    const int q = t >> 6, h = q >> 1, col = q & 1;
    const WallishGap g = sm.gaps[q];
    const int iend = g.ok ? g.b1 : G::H - 1;
    for (int i = g.b0 + (t & 63); i <= iend; i += 64) {
      double* y = reinterpret_cast<double*>(sm.X + wpos(h, i)) + col;
      *y = wallish_fill(*y, i, g);
    }
  }
  __syncthreads();
  // DST-III and exp(.)/k on the kept rows                                                   (:409-416)
  wallish_dst3_pre(t, sm.X, v, a.twd);
  fft4096(t, v, sm.S, a.tw1, a.tw2);
  // FFT bins of the DST-III, bin order (coalesced); exp(.)/k and the transposition happen in wallish_unpack_kernel
#pragma unroll
  for (int r = 0; r < 16; ++r) __stcs(zrow + t + 256 * r, v[r]);
}

// ---- layout changes around the fused kernel: both are 64-row x 32-column tile transpositions through shared memory ----
// pack: pklin [4096, ncols] -> packed [npairs, 4096] double2 = sign * log(k P), Makhoul order (n < N/2: j = 2n, else j = 2(N-1-n)+1).
// 64 consecutive rows j0 .. j0+63 hold 32 ascending samples n = j0/2 + i (even j) and 32 descending n = N-1-j0/2-i (odd j).
__global__ void __launch_bounds__(256) wallish_pack_kernel(const double* __restrict__ klin, const double* __restrict__ pklin, const long long ld,
                                                           const long long ncols, double2* __restrict__ packed) {
  typedef WallishGeo G;
  __shared__ double tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j0 = blockIdx.y * 64;
  const long long c0 = (long long)blockIdx.x * 32;
  const long long col = c0 + tx < ncols ? c0 + tx : ncols - 1;      // odd column count: the last pair repeats its column
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = j0 + ty + 8 * i;
    const double v = log(klin[j] * __ldcs(pklin + (long long)j * ld + col));
    tile[ty + 8 * i][tx] = (j & 1) ? -v : v;
  }
  __syncthreads();
  // unit u = (pair p, parity): 32 lanes write 32 consecutive samples
  for (int u = ty; u < 32; u += 8) {
    const int p = u >> 1, odd = u & 1;
    const long long pair = (c0 >> 1) + p;
    if (2 * pair >= ncols) continue;
    int jj, n;
    if (!odd) { jj = 2 * tx; n = (j0 >> 1) + tx; }
    else { jj = 2 * (31 - tx) + 1; n = G::N - 1 - (j0 >> 1) - (31 - tx); }
    packed[pair * G::N + n] = mk2(tile[jj][2 * p], tile[jj][2 * p + 1]);
  }
}

// unpack: packed [npairs, 4096] FFT bins m of the DST-III -> vals [nl + (j - i0), col] = exp(sign * bin / N) / k_j for rows
// i0 <= j < i1 (bao_filter.py:412-416); j odd <-> m = (j+1)/2, j even <-> m = N - j/2 (m = 0 for j = 0), sign as in
// wallish_dst3_out_index.
__global__ void __launch_bounds__(256) wallish_unpack_kernel(const double* __restrict__ klin, const double2* __restrict__ packed,
                                                             const long long ncols, const int i0, const int i1, const int nl,
                                                             double* __restrict__ vals) {
  typedef WallishGeo G;
  __shared__ double tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j0 = blockIdx.y * 64;
  const long long c0 = (long long)blockIdx.x * 32;
  for (int u = ty; u < 32; u += 8) {
    const int p = u >> 1, odd = u & 1;
    const long long pair = (c0 >> 1) + p;
    if (2 * pair >= ncols) continue;
    int jj, m;
    if (odd) { jj = 2 * tx + 1; m = ((j0 + jj) + 1) >> 1; }
    else { jj = 2 * (31 - tx); m = (G::N - ((j0 + jj) >> 1)) & (G::N - 1); }
    const double2 z = __ldcs(packed + pair * G::N + m);
    tile[jj][2 * p] = z.x;
    tile[jj][2 * p + 1] = z.y;
  }
  __syncthreads();
  const long long col = c0 + tx;
  if (col >= ncols) return;
  const double inv = 1. / G::N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = j0 + ty + 8 * i;
    if (j < i0 || j >= i1) continue;
    const double sign = (j & 1) ? -1. : 1.;           // j = 0 (m = 0) and even j: +1; odd j: -1
    vals[(long long)(nl + j - i0) * ncols + col] = exp(sign * inv * tile[ty + 8 * i][tx]) / klin[j];
  }
}

// rows of the knot matrix copied from the unfiltered spectrum: k < 5e-4 (first nl rows of kout) and k > 2 (last nr)
__global__ void wallish_edges_kernel(const double* __restrict__ pkout, const long long ld, const int nk, const long long ncols, const int nl,
                                     const int nmid, const int nr, double* __restrict__ vals) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int r = blockIdx.y;   // 0 .. nl+nr-1
  if (col >= ncols) return;
  const int src = r < nl ? r : nk - nr + (r - nl);
  const int dst = r < nl ? r : nl + nmid + (r - nl);
  vals[(long long)dst * ncols + col] = pkout[(long long)src * ld + col];
}

// evaluate the final clamped spline at kout[q] (interval idx[q] precomputed: the knots are shared) and blend
__global__ void wallish_final_kernel(const double* __restrict__ knots, const double* __restrict__ vals, const double* __restrict__ slopes,
                                     const long long ncols, const double* __restrict__ kout, const int* __restrict__ idx,
                                     const double* __restrict__ pkout, const long long ld, double* __restrict__ pknow) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (col >= ncols) return;
  const double k = kout[q];
  const int i = idx[q];
  if (i < 0) { pknow[(long long)q * ld + col] = nan(""); return; }
  const long long o = (long long)i * ncols + col;
  const double smooth = spline_poly(knots[i], knots[i + 1], vals[o], vals[o + ncols], slopes[o], slopes[o + ncols], k, 0);   // :420
  const double th = k > 1. ? exp(-400. * (k - 1.) * (k - 1.)) : 1.;                                                          // :425-431, scale=20
  const double pk = pkout[(long long)q * ld + col];
  const double wiggles = (pk / smooth - 1.) * th + 1.;                                                                      // :422
  pknow[(long long)q * ld + col] = pk / wiggles;                                                                          // :423
}

// ---- standalone DST-II / DST-III (orthonormal), N = 4096, along axis 0 of [4096, ncols] ------------------------------
template <int TYPE>
__global__ void __launch_bounds__(256, 1) dst_kernel(const double* __restrict__ in, double* __restrict__ out, const long long ncols,
                                                     const double2* tw1, const double2* tw2, const double2* twd) {
  typedef WallishGeo G;
  extern __shared__ double2 smem_raw[];
  const WallishSmem sm(smem_raw);
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
    dst2_in_smem(t, v, sm, tw1, tw2, twd);
    for (int kk = t; kk < G::N; kk += 256) {
      const double2 x = sm.X[wpos(kk & 1, kk >> 1)];
      double* dst = out + (long long)kk * ncols + col0;
      dst[0] = x.x;
      if (has1) dst[1] = x.y;
    }
  } else {
    for (int kk = t; kk < G::N; kk += 256) {
      const double* row = in + (long long)kk * ncols + col0;
      sm.X[wpos(kk & 1, kk >> 1)] = mk2(row[0], has1 ? row[1] : 0.);
    }
    __syncthreads();
    wallish_dst3_pre(t, sm.X, v, twd);
    fft4096(t, v, sm.S, tw1, tw2);
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

int cpf_wallish2018(const double* klin, const double* pklin, int nlin, const double* kout, const double* pkout, int nk,
                    int64_t ncols, double* pknow, int32_t* boxes, int on_device, int device, void* stream_) {
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

  // the two wavenumber grids are needed on the host (knot selection, interval search) and on the device
  std::vector<double> h_klin(nlin), h_kout(nk);
  if (on_device) {
    CPF_CUDA(cudaMemcpyAsync(h_klin.data(), klin, nlin * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaMemcpyAsync(h_kout.data(), kout, nk * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  } else {
    h_klin.assign(klin, klin + nlin);
    h_kout.assign(kout, kout + nk);
  }
  for (int i = 1; i < nlin; ++i) if (!(h_klin[i] > h_klin[i - 1])) return fail(CPF_EINVAL, "cpf_wallish2018: klin must be strictly increasing");
  for (int i = 1; i < nk; ++i) if (!(h_kout[i] > h_kout[i - 1])) return fail(CPF_EINVAL, "cpf_wallish2018: kout must be strictly increasing");
  int i0 = 0, i1 = nlin, nl = 0, nr = 0;
  while (i0 < nlin && !(h_klin[i0] > 1e-2)) ++i0;                    // mask = (k > 1e-2) & (k < 1.5)   (:415)
  while (i1 > i0 && !(h_klin[i1 - 1] < 1.5)) --i1;
  while (nl < nk && h_kout[nl] < 5e-4) ++nl;                         // mask_left = self.k < 5e-4        (:417)
  while (nr < nk - nl && h_kout[nk - 1 - nr] > 2.) ++nr;             // mask_right = self.k > 2
  const int nmid = i1 - i0, nknots = nl + nmid + nr;
  if (nknots < 2) return fail(CPF_EINVAL, "cpf_wallish2018: fewer than two knots survive the k cuts");
  std::vector<double> h_knots(nknots);
  for (int i = 0; i < nl; ++i) h_knots[i] = h_kout[i];
  for (int i = 0; i < nmid; ++i) h_knots[nl + i] = h_klin[i0 + i];
  for (int i = 0; i < nr; ++i) h_knots[nl + nmid + i] = h_kout[nk - nr + i];
  for (int i = 1; i < nknots; ++i) if (!(h_knots[i] > h_knots[i - 1])) return fail(CPF_EINVAL, "cpf_wallish2018: spliced knots are not increasing");
  std::vector<int> h_idx(nk);
  for (int q = 0; q < nk; ++q) {
    // outside the spliced knots the reference's CubicSpline(extrapolate=False) yields NaN (:420), hence pknow = NaN there (:422-423)
    const bool outside = h_kout[q] < h_knots[0] || h_kout[q] > h_knots[nknots - 1];
    h_idx[q] = outside ? -1 : spline_interval(h_knots.data(), nknots, h_kout[q]);
  }

  const size_t lin_bytes = (size_t)nlin * ncols * sizeof(double), out_bytes = (size_t)nk * ncols * sizeof(double);
  // columns are processed in chunks so that the work arrays (packed: 64 KB, knot values + slopes: 59 KB per spectrum) stay within the
  // private scratch pool's release threshold whatever the batch
  const long long chunk = ncols < 8192 ? ncols : 8192;
  const size_t knot_bytes = (size_t)nknots * chunk * sizeof(double);
  ScratchBuf d_klin, d_pklin, d_kout, d_pkout, d_pknow, d_boxes, d_knots, d_idx, d_vals, d_slopes, d_fac, d_packed;
  const double *p_klin = klin, *p_pklin = pklin, *p_kout = kout, *p_pkout = pkout;
  double* p_pknow = pknow;
  int* p_boxes = boxes;
  if (!on_device) {
    CPF_CUDA(d_klin.alloc(nlin * sizeof(double), stream));
    CPF_CUDA(d_pklin.alloc(lin_bytes, stream));
    CPF_CUDA(d_kout.alloc(nk * sizeof(double), stream));
    CPF_CUDA(d_pkout.alloc(out_bytes, stream));
    CPF_CUDA(d_pknow.alloc(out_bytes, stream));
    CPF_CUDA(cudaMemcpyAsync(d_klin.p, klin, nlin * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(d_pklin.p, pklin, lin_bytes, cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(d_kout.p, kout, nk * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(d_pkout.p, pkout, out_bytes, cudaMemcpyHostToDevice, stream));
    p_klin = (const double*)d_klin.p; p_pklin = (const double*)d_pklin.p; p_kout = (const double*)d_kout.p; p_pkout = (const double*)d_pkout.p;
    p_pknow = (double*)d_pknow.p;
    if (boxes) {
      CPF_CUDA(d_boxes.alloc((size_t)ncols * 4 * sizeof(int), stream));
      p_boxes = (int*)d_boxes.p;
    }
  }
  CPF_CUDA(d_knots.alloc(nknots * sizeof(double), stream));
  CPF_CUDA(d_idx.alloc(nk * sizeof(int), stream));
  CPF_CUDA(d_vals.alloc(knot_bytes, stream));
  CPF_CUDA(d_slopes.alloc(knot_bytes, stream));
  CPF_CUDA(d_fac.alloc(4 * (size_t)nknots * sizeof(double), stream));
  CPF_CUDA(d_packed.alloc((size_t)((chunk + 1) / 2) * WallishGeo::N * sizeof(double2), stream));
  // the factors of the final clamped spline depend on the spliced knots only: computed here, on the host copy
  std::vector<double> h_fac(4 * (size_t)nknots);
  spline_factor_host(h_knots.data(), nknots, 1, h_fac.data());
  CPF_CUDA(cudaMemcpyAsync(d_fac.p, h_fac.data(), h_fac.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  CPF_CUDA(cudaMemcpyAsync(d_knots.p, h_knots.data(), nknots * sizeof(double), cudaMemcpyHostToDevice, stream));
  CPF_CUDA(cudaMemcpyAsync(d_idx.p, h_idx.data(), nk * sizeof(int), cudaMemcpyHostToDevice, stream));
  CPF_CUDA(cudaFuncSetAttribute(wallish_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWallishSmemBytes));
  for (long long c0 = 0; c0 < ncols; c0 += chunk) {
    const long long cc = ncols - c0 < chunk ? ncols - c0 : chunk;
    WallishArgs a;
    a.klin = p_klin; a.pklin = p_pklin + c0; a.packed = (double2*)d_packed.p; a.ncols = cc; a.ld = ncols; a.i0 = i0; a.i1 = i1; a.nl = nl;
    a.vals = (double*)d_vals.p; a.boxes = p_boxes ? p_boxes + 4 * c0 : nullptr;
    a.tw1 = wt.tw1; a.tw2 = wt.tw2; a.twd = wt.twd; a.wtab = wt.wtab;
    const dim3 tgrid((unsigned)((cc + 31) / 32), WallishGeo::N / 64);
    wallish_pack_kernel<<<tgrid, 256, 0, stream>>>(p_klin, a.pklin, ncols, cc, a.packed);
    wallish_fused_kernel<<<(unsigned)((cc + 1) / 2), 256, kWallishSmemBytes, stream>>>(a);
    wallish_unpack_kernel<<<tgrid, 256, 0, stream>>>(p_klin, a.packed, cc, i0, i1, nl, a.vals);
    CPF_CUDA(cudaGetLastError());
    const unsigned ctile = (unsigned)((cc + 127) / 128);
    if (nl + nr > 0) {
      wallish_edges_kernel<<<dim3(ctile, (unsigned)(nl + nr)), 128, 0, stream>>>(p_pkout + c0, ncols, nk, cc, nl, nmid, nr, (double*)d_vals.p);
      CPF_CUDA(cudaGetLastError());
    }
    CPF_TRY(spline_fit_device((const double*)d_knots.p, (const double*)d_vals.p, nknots, cc, 1, (double*)d_slopes.p, (double*)d_fac.p, stream, true));
    wallish_final_kernel<<<dim3(ctile, (unsigned)nk), 128, 0, stream>>>((const double*)d_knots.p, (const double*)d_vals.p, (const double*)d_slopes.p,
                                                                          cc, p_kout, (const int*)d_idx.p, p_pkout + c0, ncols, p_pknow + c0);
    CPF_CUDA(cudaGetLastError());
  }
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(pknow, p_pknow, out_bytes, cudaMemcpyDeviceToHost, stream));
    if (boxes) CPF_CUDA(cudaMemcpyAsync(boxes, p_boxes, (size_t)ncols * 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  }
  // host vectors (knots, idx) must outlive the async uploads; scratch is stream-ordered
  CPF_CUDA(cudaStreamSynchronize(stream));
  return CPF_OK;
}

}  // extern "C"
