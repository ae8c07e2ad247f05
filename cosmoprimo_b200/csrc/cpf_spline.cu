// cpf_spline.cu — batched cubic splines on a shared abscissa grid: fit (natural / clamped) and evaluation of the
// value or a derivative.  Replaces scipy.interpolate.CubicSpline + PPoly.__call__ as used by
// cosmoprimo/jax.py:139-196 (Interpolator1D, numpy path) and cosmoprimo/bao_filter.py:377-382, 400-402, 420.
//
// Layout is the reference's: knots along axis 0, one column per spline, y[nx, ncols] row-major.  One thread owns one
// column, so every load/store of a warp is a contiguous 256-byte run.  HBM-bound: the fit streams y once and the slope
// array twice (forward elimination writes the reduced right-hand side into it, back substitution rewrites it).
//
// Mathematics (same system scipy assembles, scipy/interpolate/_cubic.py): slopes s_i = y'(x_i) solve
//     dx_i s_{i-1} + 2 (dx_{i-1} + dx_i) s_i + dx_{i-1} s_{i+1} = 3 (dx_i m_{i-1} + dx_{i-1} m_i),  m_i = (y_{i+1}-y_i)/dx_i
// with end rows   natural: 2 dx_0 s_0 + dx_0 s_1 = 3 (y_1 - y_0)   |   clamped: s_0 = 0   (mirrored at the other end).
// The matrix depends on x only, so its LU factors are computed once per grid and shared by all columns; it is strictly
// diagonally dominant by rows, so elimination without pivoting is stable (LAPACK's dgtsv, which scipy calls, only
// pivots when dominance fails).
#include <string.h>
#include <vector>

#include "cpf_common.h"
#include "cpf_spline_core.h"
#include "cpf_fastmath.h"

namespace cpf {

// ---- factorisation: one thread, O(nx) ------------------------------------------------------------------------------
// fac[0*nx + i] = Lw_i, fac[1*nx + i] = cp_i, fac[2*nx + i] = P_i, fac[3*nx + i] = Q_i (spline_factor_step)
__global__ void spline_factor_kernel(const double* __restrict__ x, const int nx, const int bc, double* __restrict__ fac) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double cprev = 0.;
  for (int i = 0; i < nx; ++i) {
    double Lw, P, Q;
    spline_factor_step(x, nx, bc, i, cprev, Lw, P, Q);
    fac[i] = Lw;
    fac[nx + i] = cprev;
    fac[2 * nx + i] = P;
    fac[3 * nx + i] = Q;
  }
}

// the same on the host, for callers that hold the knots there (cpf_wallish2018): 4*nx doubles
void spline_factor_host(const double* x, const int nx, const int bc, double* fac) {
  double cprev = 0.;
  for (int i = 0; i < nx; ++i) {
    double Lw, P, Q;
    spline_factor_step(x, nx, bc, i, cprev, Lw, P, Q);
    fac[i] = Lw;
    fac[nx + i] = cprev;
    fac[2 * nx + i] = P;
    fac[3 * nx + i] = Q;
  }
}

// optional log10 of the abscissae / ordinates (Interpolator1D's interp_x='log' / interp_fun='log', jax.py:152-153)
__global__ void log10_kernel(const double* __restrict__ in, double* __restrict__ out, const long long count, const int apply) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < count) out[i] = apply ? fast_log10(in[i]) : in[i];
}


// ---- log-log tables with power-law continuation knots (interpolator.py:42-87 `_pad_log`, 343-351; NaN rules of jax.py:161-172) ----
// y [nx, ncols] tabulated spectra -> ly [nx + 4, ncols] = log10(y) in rows 2 .. nx + 1 and, in rows 0, 1, nx + 2, nx + 3, the straight
// lines through the two lowest / two highest samples evaluated at the continuation knots lx[0], lx[1], lx[nx + 2], lx[nx + 3] (lx =
// log10 of the padded knots).  nan_count[col] += number of NaN cells of the column (log10 of a negative sample is NaN; of zero, -inf,
// which the reference lets through as well).  One thread per (column, chunk of PADLOG_ROWS rows): a warp touches 256 contiguous bytes.
#define PADLOG_ROWS 64
__global__ void __launch_bounds__(128) padlog_kernel(const double* __restrict__ y, const double* __restrict__ lx, const int nx,
                                                     const long long ncols, double* __restrict__ ly, int* __restrict__ nan_count) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int r0 = blockIdx.y * PADLOG_ROWS, r1 = min(nx, r0 + PADLOG_ROWS);
  int bad = 0;
#pragma unroll 4
  for (int i = r0; i < r1; ++i) {
    const double v = fast_log10(__ldcs(y + (long long)i * ncols + col));
    bad += isnan(v) ? 1 : 0;
    ly[(long long)(i + 2) * ncols + col] = v;
  }
  if (blockIdx.y == 0) {                         // continuation below the table: slope of the two lowest samples
    const double a = fast_log10(y[col]), b = fast_log10(y[ncols + col]);
    const double x0 = lx[2], slope = (b - a) / (lx[3] - x0);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double v = a + slope * (lx[j] - x0);
      bad += isnan(v) ? 1 : 0;
      ly[(long long)j * ncols + col] = v;
    }
  }
  if (blockIdx.y == gridDim.y - 1) {             // continuation above the table: slope of the two highest samples
    const double a = fast_log10(y[(long long)(nx - 2) * ncols + col]), b = fast_log10(y[(long long)(nx - 1) * ncols + col]);
    const double x1 = lx[nx + 1], slope = (b - a) / (x1 - lx[nx]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double v = b + slope * (lx[nx + 2 + j] - x1);
      bad += isnan(v) ? 1 : 0;
      ly[(long long)(nx + 2 + j) * ncols + col] = v;
    }
  }
  if (bad) atomicAdd(nan_count + col, bad);
}

// NaN screening of a table (jax.py:161-172): nan_count[col] += number of NaN (or, for a log10 ordinate, negative) cells of the column
__global__ void __launch_bounds__(128) nan_count_kernel(const double* __restrict__ y, const int nx, const long long ncols, const int neg_is_nan,
                                                        int* __restrict__ nan_count) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int r0 = blockIdx.y * PADLOG_ROWS, r1 = min(nx, r0 + PADLOG_ROWS);
  int bad = 0;
#pragma unroll 8
  for (int i = r0; i < r1; ++i) {
    const double v = y[(long long)i * ncols + col];
    bad += (isnan(v) || (neg_is_nan && v < 0.)) ? 1 : 0;
  }
  if (bad) atomicAdd(nan_count + col, bad);
}

// ---- per-column forward elimination + back substitution ------------------------------------------------------------
// One thread per (column, chunk of SPLINE_CHUNK knots).  The recurrences are one FMA deep per knot; ordinates are loaded a
// block of SPLINE_U knots ahead so that the HBM latency is paid once per block, not once per knot.  Chunks make enough
// threads to hide that latency when there are few columns: the system is strictly diagonally dominant, a perturbation of
// the recurrence decays by >= 2 (typically 3.7) per knot, so a chunk starts SPLINE_WARM knots early from zero (forward) /
// from the reduced right-hand side (backward) and is exact to < 1e-19 where it starts storing.  nx <= SPLINE_SERIAL_MAX:
// one chunk, i.e. the plain Thomas algorithm.
#define SPLINE_U 16
#define SPLINE_CHUNK 256
#define SPLINE_WARM 64
#define SPLINE_SERIAL_MAX 512

// forward: dp_i = P_i (y_i - y_{i-1}) + Q_i (y_{i+1} - y_i) - Lw_i dp_{i-1}
__global__ void __launch_bounds__(128) spline_forward_kernel(const double* __restrict__ y, const double* __restrict__ fac, const int nx,
                                                              const long long ncols, const int chunk, const int nak,
                                                              double* __restrict__ dp) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int first = blockIdx.y * chunk, last = min(nx, first + chunk) - 1;
  const int start = max(0, first - SPLINE_WARM);
  const double* Lw = fac;
  const double* P = fac + 2 * nx;
  const double* Q = fac + 3 * nx;
  const double* yc = y + col;
  double* dc = dp + col;
  double nxt[SPLINE_U];
#pragma unroll
  for (int u = 0; u < SPLINE_U; ++u) nxt[u] = (start + 1 + u < nx) ? yc[(long long)(start + 1 + u) * ncols] : 0.;
  double ym = start > 0 ? yc[(long long)(start - 1) * ncols] : 0., y0 = yc[(long long)start * ncols], dprev = 0.;
  double ymm = start > 1 ? yc[(long long)(start - 2) * ncols] : 0.;     // only the not-a-knot end row looks two knots back
  for (int base = start; base <= last; base += SPLINE_U) {
    double cur[SPLINE_U];
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) cur[u] = nxt[u];
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) {
      const int j = base + SPLINE_U + 1 + u;
      nxt[u] = (j < nx && j <= last + 1) ? yc[(long long)j * ncols] : 0.;
    }
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) {
      const int i = base + u;
      if (i <= last) {
        const double yp = cur[u];
        double a = y0 - ym, b = yp - y0;
        if (nak) {                                          // not-a-knot end rows use the two intervals next to the end
          if (i == 0) { a = yp - y0; b = cur[u + 1 < SPLINE_U ? u + 1 : u] - yp; }      // i = 0 is always u = 0 of the first block
          else if (i == nx - 1) { a = ym - ymm; b = y0 - ym; }
        }
        const double t = fma(__ldg(P + i), a, __ldg(Q + i) * b);
        dprev = fma(-__ldg(Lw + i), dprev, t);
        if (i >= first) dc[(long long)i * ncols] = dprev;
        ymm = ym;
        ym = y0;
        y0 = yp;
      }
    }
  }
}

// backward: s_i = dp_i - cp_i s_{i+1}
__global__ void __launch_bounds__(128) spline_backward_kernel(const double* __restrict__ dp, const double* __restrict__ fac, const int nx,
                                                               const long long ncols, const int chunk, double* __restrict__ s) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int first = blockIdx.y * chunk, last = min(nx, first + chunk) - 1;
  const int end = min(nx - 1, last + SPLINE_WARM);
  const double* cp = fac + nx;
  const double* dc = dp + col;
  double* sc = s + col;
  double snext = dc[(long long)end * ncols];
  if (end == last) sc[(long long)end * ncols] = snext;      // last knot of the spline: s = dp
  int top = end - 1;
  double nxt[SPLINE_U];
#pragma unroll
  for (int u = 0; u < SPLINE_U; ++u) nxt[u] = (top - u >= first) ? dc[(long long)(top - u) * ncols] : 0.;
  for (; top >= first; top -= SPLINE_U) {
    double cur[SPLINE_U];
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) cur[u] = nxt[u];
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) {
      const int j = top - SPLINE_U - u;
      nxt[u] = j >= first ? dc[(long long)j * ncols] : 0.;
    }
#pragma unroll
    for (int u = 0; u < SPLINE_U; ++u) {
      const int i = top - u;
      if (i >= first) {
        snext = fma(-__ldg(cp + i), snext, cur[u]);
        if (i <= last) sc[(long long)i * ncols] = snext;
      }
    }
  }
}

// ---- evaluation ----------------------------------------------------------------------------------------------------
// Everything that depends on the query only -- log10 of the abscissa, bounds test, interval search -- is done once per query
// by spline_query_kernel (it used to be repeated by every thread: ~3/4 of the instructions of an evaluation);
// qi[q] = interval index, or -1 when the result is NaN; qx[q] = offset of the (transformed) abscissa from the left knot of its interval,
// qx[nq + q] = 1 / width of the interval (the one division of spline_poly: it used to be taken per column and query).
__global__ void spline_query_kernel(const double* __restrict__ x, const int nx, const double* __restrict__ xq, const int nq,
                                    const int extrap, const int log_x, const double xmin_raw, const double xmax_raw,
                                    double* __restrict__ qx, int* __restrict__ qi) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const double raw = xq[q];
  // bounds are tested on the raw abscissa, before the log10 (jax.py:188-189)
  const bool inside = raw >= xmin_raw && raw <= xmax_raw;
  const double xv = log_x ? log10(raw) : raw;
  const int i = ((!inside && !extrap) || !(xv == xv)) ? -1 : spline_interval(x, nx, xv);
  qi[q] = i;
  qx[q] = i < 0 ? 0. : xv - x[i];
  qx[nq + q] = i < 0 ? 0. : 1. / (x[i + 1] - x[i]);
}

// out[q, col]: block = (column tile, SPLINE_EVAL_Q consecutive queries).  A thread keeps the knot values and slopes of its column while
// consecutive queries fall into the same interval (sorted query grids: the usual case), so the table is read about once instead of once
// per query, and one block writes SPLINE_EVAL_Q x 1 KB instead of 1 KB (it used to be one tiny block per query: 0.6 TB/s).
#define SPLINE_EVAL_Q 16
__global__ void __launch_bounds__(128) spline_eval_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                           const double* __restrict__ s, const long long ncols,
                                                           const double* __restrict__ qx, const int* __restrict__ qi, const int nq, const int nu,
                                                           const int log_y, double* __restrict__ out) {
  const int q0 = blockIdx.y * SPLINE_EVAL_Q, q1 = min(nq, q0 + SPLINE_EVAL_Q);
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  int have = -2;
  double y1 = 0., s1 = 0.;
  SplineCubic cub = {0., 0., 0., 0.};
  for (int q = q0; q < q1; ++q) {
    const int i = __ldg(qi + q);
    double r;
    if (i < 0) {
      r = nan("");
    } else {
      if (i != have) {                                               // coefficients once per (interval, column)
        const long long o = (long long)i * ncols + col;
        double y0, s0;
        if (i == have + 1) { y0 = y1; s0 = s1; }                     // the next interval shares a knot
        else { y0 = y[o]; s0 = s[o]; }
        y1 = y[o + ncols]; s1 = s[o + ncols];
        cub = spline_coeffs(__ldg(qx + nq + q), y0, y1, s0, s1);
        have = i;
      }
      r = spline_cubic_eval(cub, __ldg(qx + q), nu);
      if (log_y) r = fast_exp10(r);   // 10**tmp, jax.py:191
    }
    __stcs(out + (long long)q * ncols + col, r);
  }
}

// transposed evaluation: out[col, q] (rows = splines: the layout cpf_fftlog reads), 32 x 32 tiles through shared memory so
// that both the knot-matrix reads (contiguous in col) and the stores (contiguous in q) are coalesced
#define SPLINE_EVALT_Q 128          // queries per block: 8 groups of 16 consecutive queries; 32 columns per block
__global__ void __launch_bounds__(256) spline_eval_t_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                            const double* __restrict__ s, const long long ncols,
                                                            const double* __restrict__ qx, const int* __restrict__ qi, const int nq,
                                                            const int nu, const int log_y, double* __restrict__ out) {
  __shared__ double tile[SPLINE_EVALT_Q][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long c0 = (long long)blockIdx.x * 32;
  const int q0 = blockIdx.y * SPLINE_EVALT_Q;
  const long long col = c0 + tx;
  // a thread takes 16 consecutive queries of its column and keeps the knot values / slopes while they stay in one interval
  int have = -2;
  double y1 = 0., s1 = 0.;
  SplineCubic cub = {0., 0., 0., 0.};
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const int qq = 16 * ty + j, q = q0 + qq;
    double r = 0.;
    if (q < nq && col < ncols) {
      const int i = __ldg(qi + q);
      if (i < 0) {
        r = nan("");
      } else {
        if (i != have) {
          const long long o = (long long)i * ncols + col;
          double y0, s0;
          if (i == have + 1) { y0 = y1; s0 = s1; }
          else { y0 = y[o]; s0 = s[o]; }
          y1 = y[o + ncols]; s1 = s[o + ncols];
          cub = spline_coeffs(__ldg(qx + nq + q), y0, y1, s0, s1);
          have = i;
        }
        r = spline_cubic_eval(cub, __ldg(qx + q), nu);
        if (log_y) r = fast_exp10(r);
      }
    }
    tile[qq][tx] = r;
  }
  __syncthreads();
  // one warp per column: 128 consecutive queries = four 256-byte runs of a row of the result
  for (int cc = ty; cc < 32; cc += 8) {
    const long long c = c0 + cc;
    if (c >= ncols) continue;
#pragma unroll
    for (int k = 0; k < SPLINE_EVALT_Q / 32; ++k) {
      const int q = q0 + 32 * k + tx;
      if (q < nq) __stcs(out + c * nq + q, tile[32 * k + tx][cc]);
    }
  }
}

// ---- row layout: splines along the LAST axis, windowed weights (cpf_spline_core.h) ------------------------------------
// weights kernel: one warp per query (the Thomas sweep is serial on lane 0, in shared memory when the window fits; the
// weights and their trimming use the whole warp).  wq [nq, LW] weights, meta [3, nq] = (first knot, length or -1 for NaN),
// offset of the first kept weight
__global__ void __launch_bounds__(32) spline_row_weights_kernel(const double* __restrict__ x, const int nx, const int bc, const int W,
                                                               const int LW, const double* __restrict__ xq, const int nq,
                                                               const int extrap, const int use_smem, double* __restrict__ wq,
                                                               double* __restrict__ work, int* __restrict__ meta) {
  extern __shared__ double sw_smem[];
  const int q = blockIdx.x, lane = threadIdx.x;
  const double xv = xq[q];
  double* w = wq + (size_t)q * LW;
  const bool inside = xv >= x[0] && xv <= x[nx - 1];
  if ((!inside && !extrap) || !(xv == xv)) {
    if (lane == 0) { meta[2 * q] = 0; meta[2 * q + 1] = -1; meta[2 * nq + q] = 0; }
    return;
  }
  const SplineWindow sw = spline_window_setup(x, nx, bc, W, xv);
  double* cp = use_smem ? sw_smem : work + (size_t)q * 2 * LW;
  double* z = cp + LW;
  if (lane == 0) spline_window_solve(x + sw.a, sw, cp, z);
  __syncwarp();
  double wmax = 0.;
  for (int j = lane; j < sw.L; j += 32) {
    const double wj = spline_window_weight(x + sw.a, sw, z, j);
    w[j] = wj;
    wmax = fmax(wmax, fabs(wj));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  __syncwarp();
  // trimmed range (spline_trim_weights): first / last weight above 1e-40 of the largest
  const double thr = 1e-40 * wmax;
  int lo = sw.L, hi = -1;
  for (int j = lane; j < sw.L; j += 32)
    if (fabs(w[j]) > thr) { lo = min(lo, j); hi = max(hi, j); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (hi < lo) { lo = 0; hi = 0; }
  if (lane == 0) {
    meta[2 * q] = sw.a + lo;
    meta[2 * q + 1] = hi - lo + 1;
    meta[2 * nq + q] = lo;
  }
}

// dot kernel: one warp per row; out[q, row] = sum_j w[q, j] y[row, first_q + j].  Window loads are contiguous runs.
__global__ void __launch_bounds__(256) spline_rows_dot_kernel(const double* __restrict__ y, const int nx, const long long rows,
                                                              const double* __restrict__ wq, const int* __restrict__ meta,
                                                              const int nq, const int LW, const int post_sqrt,
                                                              double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* yr = y + row * nx;
  for (int q = 0; q < nq; ++q) {
    const int first = meta[2 * q], L = meta[2 * q + 1];
    double acc = 0.;
    if (L > 0) {
      const double* w = wq + (size_t)q * LW + meta[2 * nq + q];
      for (int j = lane; j < L; j += 32) acc = fma(w[j], yr[first + j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    } else {
      acc = nan("");
    }
    if (lane == (q & 31)) out[(long long)q * rows + row] = post_sqrt ? sqrt(acc) : acc;     // sigma(r) from sigma^2(r), interpolator.py:573
  }
}

// device-resident fit, shared with cpf_wallish.cu: slopes s[nx, ncols] of the splines through y[nx, ncols] on the
// knots x[nx]; fac is scratch of 4*nx doubles (already filled by spline_factor_host when fac_ready)
int spline_fit_device(const double* d_x, const double* d_y, int nx, long long ncols, int bc, double* d_s, double* d_fac,
                      cudaStream_t stream, bool fac_ready) {
  if (!fac_ready) spline_factor_kernel<<<1, 32, 0, stream>>>(d_x, nx, bc, d_fac);
  if (ncols > 0) {
    ScratchBuf dp;
    CPF_CUDA(dp.alloc((size_t)nx * (size_t)ncols * sizeof(double), stream));
    const int chunk = nx <= SPLINE_SERIAL_MAX ? nx : SPLINE_CHUNK;
    const dim3 grid((unsigned)((ncols + 127) / 128), (unsigned)((nx + chunk - 1) / chunk));
    spline_forward_kernel<<<grid, 128, 0, stream>>>(d_y, d_fac, nx, ncols, chunk, bc == 2 ? 1 : 0, (double*)dp.p);
    spline_backward_kernel<<<grid, 128, 0, stream>>>((const double*)dp.p, d_fac, nx, ncols, chunk, d_s);
  }
  CPF_CUDA(cudaGetLastError());
  return CPF_OK;
}

}  // namespace cpf

using namespace cpf;

struct cpf_spline {
  int nx, bc, log_x, log_y, extrap, device;
  long long ncols;
  double xmin_raw, xmax_raw;
  double* d_x = nullptr;   // (log10 of) abscissae
  double* d_y = nullptr;   // (log10 of) ordinates [nx, ncols]
  double* d_s = nullptr;   // slopes [nx, ncols]
  // the three arrays come from the library's private stream-ordered pool (no device-wide synchronisation at create / destroy, unlike
  // cudaMalloc / cudaFree); they are released in stream order behind the last stream that used the spline
  mutable cudaStream_t last_stream = nullptr;
};

extern "C" {

int cpf_spline_destroy(cpf_spline* sp) {
  if (!sp) return CPF_OK;
  DeviceGuard guard(sp->device);
  if (sp->d_x) cudaFreeAsync(sp->d_x, sp->last_stream);
  if (sp->d_y) cudaFreeAsync(sp->d_y, sp->last_stream);
  if (sp->d_s) cudaFreeAsync(sp->d_s, sp->last_stream);
  delete sp;
  return CPF_OK;
}

int cpf_spline_create(cpf_spline** out, const double* x, const double* y, int nx, int64_t ncols, int bc, int log_x, int log_y,
                      int extrap, int on_device, int device, void* stream_) {
  if (!out) return fail(CPF_EINVAL, "cpf_spline_create: null handle pointer");
  *out = nullptr;
  if (!x || (!y && ncols > 0)) return fail(CPF_EINVAL, "cpf_spline_create: null buffer");
  if (nx < 2) return fail(CPF_EINVAL, "cpf_spline_create: need at least 2 knots, got %d", nx);
  if (ncols < 0) return fail(CPF_EINVAL, "cpf_spline_create: negative column count");
  if (bc < 0 || bc > 2) return fail(CPF_EINVAL, "cpf_spline_create: bc must be 0 (natural), 1 (clamped) or 2 (not-a-knot)");
  if (bc == 2 && nx < 4) return fail(CPF_EUNSUPPORTED, "cpf_spline_create: not-a-knot ends need at least 4 knots, got %d", nx);
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_spline_create: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  cpf_spline* sp = new cpf_spline();
  sp->nx = nx; sp->bc = bc; sp->log_x = log_x ? 1 : 0; sp->log_y = log_y ? 1 : 0; sp->extrap = extrap ? 1 : 0;
  sp->device = device; sp->ncols = ncols;
  const size_t cells = (size_t)nx * (size_t)ncols;
  int rc = CPF_OK;
  ScratchBuf fac, stage_x, stage_y;
  do {
    cudaError_t e;
#define SP_CUDA(call) if ((e = (call)) != cudaSuccess) { rc = fail(CPF_ECUDA, "%s: %s", #call, cudaGetErrorString(e)); break; }
    sp->last_stream = stream;
    cudaMemPool_t pool = scratch_pool(device);
    if (!pool) { rc = fail(CPF_ECUDA, "cpf_spline_create: no memory pool on device %d", device); break; }
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_x, nx * sizeof(double), pool, stream));
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_y, (cells ? cells : 1) * sizeof(double), pool, stream));
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_s, (cells ? cells : 1) * sizeof(double), pool, stream));
    // knots on the host: their (optional) logarithms and the elimination factors are O(nx) serial work that took a single device thread
    // longer than the whole fit; one small upload instead
    std::vector<double> hx((size_t)nx), tab(5 * (size_t)nx);
    if (on_device) {
      SP_CUDA(cudaMemcpyAsync(hx.data(), x, nx * sizeof(double), cudaMemcpyDeviceToHost, stream));
      SP_CUDA(cudaStreamSynchronize(stream));
    } else {
      memcpy(hx.data(), x, nx * sizeof(double));
    }
    sp->xmin_raw = hx[0]; sp->xmax_raw = hx[nx - 1];
    for (int i = 0; i < nx; ++i) tab[i] = sp->log_x ? log10(hx[i]) : hx[i];
    spline_factor_host(tab.data(), nx, bc, tab.data() + nx);
    SP_CUDA(fac.alloc(4 * (size_t)nx * sizeof(double), stream));
    SP_CUDA(cudaMemcpyAsync(sp->d_x, tab.data(), nx * sizeof(double), cudaMemcpyHostToDevice, stream));
    SP_CUDA(cudaMemcpyAsync(fac.p, tab.data() + nx, 4 * (size_t)nx * sizeof(double), cudaMemcpyHostToDevice, stream));
    const double* src_y = y;
    if (!on_device && cells) {
      SP_CUDA(stage_y.alloc(cells * sizeof(double), stream));
      SP_CUDA(cudaMemcpyAsync(stage_y.p, y, cells * sizeof(double), cudaMemcpyHostToDevice, stream));
      src_y = (const double*)stage_y.p;
    }
    if (cells) log10_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(src_y, sp->d_y, (long long)cells, sp->log_y);
    if ((rc = spline_fit_device(sp->d_x, sp->d_y, nx, ncols, bc, sp->d_s, (double*)fac.p, stream, true)) != CPF_OK) break;
    if (!on_device) SP_CUDA(cudaStreamSynchronize(stream));   // staging buffers of the caller may go away
#undef SP_CUDA
  } while (0);
  if (rc != CPF_OK) {
    cpf_spline_destroy(sp);
    return rc;
  }
  *out = sp;
  return CPF_OK;
}


int cpf_spline_create_padlog(cpf_spline** out, const double* x_padded, const double* y, int nx, int64_t ncols, int extrap,
                             uint8_t* col_flags, int on_device, int device, void* stream_) {
  if (!out) return fail(CPF_EINVAL, "cpf_spline_create_padlog: null handle pointer");
  *out = nullptr;
  if (!x_padded || (!y && ncols > 0) || (!col_flags && ncols > 0)) return fail(CPF_EINVAL, "cpf_spline_create_padlog: null buffer");
  if (nx < 2) return fail(CPF_EINVAL, "cpf_spline_create_padlog: need at least 2 tabulated knots, got %d", nx);
  if (ncols < 0) return fail(CPF_EINVAL, "cpf_spline_create_padlog: negative column count");
  const int np = nx + 4;
  for (int i = 0; i < np; ++i)
    if (!(x_padded[i] > 0.) || (i && !(x_padded[i] > x_padded[i - 1])))
      return fail(CPF_EINVAL, "cpf_spline_create_padlog: the padded knots must be positive and strictly increasing (knot %d)", i);
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_spline_create_padlog: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  cpf_spline* sp = new cpf_spline();
  sp->nx = np; sp->bc = 0; sp->log_x = 1; sp->log_y = 1; sp->extrap = extrap ? 1 : 0;
  sp->device = device; sp->ncols = ncols;
  sp->xmin_raw = x_padded[0]; sp->xmax_raw = x_padded[np - 1];
  const size_t cells = (size_t)np * (size_t)ncols, in_bytes = (size_t)nx * (size_t)ncols * sizeof(double);
  int rc = CPF_OK;
  ScratchBuf fac, stage_x, stage_y, counts;
  std::vector<int> h_counts((size_t)ncols);
  do {
    cudaError_t e;
#define SP_CUDA(call) if ((e = (call)) != cudaSuccess) { rc = fail(CPF_ECUDA, "%s: %s", #call, cudaGetErrorString(e)); break; }
    sp->last_stream = stream;
    cudaMemPool_t pool = scratch_pool(device);
    if (!pool) { rc = fail(CPF_ECUDA, "cpf_spline_create_padlog: no memory pool on device %d", device); break; }
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_x, np * sizeof(double), pool, stream));
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_y, (cells ? cells : 1) * sizeof(double), pool, stream));
    SP_CUDA(cudaMallocFromPoolAsync((void**)&sp->d_s, (cells ? cells : 1) * sizeof(double), pool, stream));
    std::vector<double> tab(5 * (size_t)np);
    for (int i = 0; i < np; ++i) tab[i] = log10(x_padded[i]);
    spline_factor_host(tab.data(), np, 0, tab.data() + np);
    SP_CUDA(fac.alloc(4 * (size_t)np * sizeof(double), stream));
    SP_CUDA(cudaMemcpyAsync(sp->d_x, tab.data(), np * sizeof(double), cudaMemcpyHostToDevice, stream));
    SP_CUDA(cudaMemcpyAsync(fac.p, tab.data() + np, 4 * (size_t)np * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (ncols > 0) {
      const double* src_y = y;
      if (!on_device) {
        SP_CUDA(stage_y.alloc(in_bytes, stream));
        SP_CUDA(cudaMemcpyAsync(stage_y.p, y, in_bytes, cudaMemcpyHostToDevice, stream));
        src_y = (const double*)stage_y.p;
      }
      SP_CUDA(counts.alloc((size_t)ncols * sizeof(int), stream));
      SP_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)ncols * sizeof(int), stream));
      const dim3 grid((unsigned)((ncols + 127) / 128), (unsigned)((nx + PADLOG_ROWS - 1) / PADLOG_ROWS));
      padlog_kernel<<<grid, 128, 0, stream>>>(src_y, sp->d_x, nx, ncols, sp->d_y, (int*)counts.p);
    }
    // NaN columns propagate through their own (independent) solves, so the fit does not wait for the flags; the copy of the counts into
    // pageable host memory blocks the host, so it is queued behind the fit kernels (it used to sit between the logarithms and the fit and
    // left the device idle while the host came back to launch them)
    if ((rc = spline_fit_device(sp->d_x, sp->d_y, np, ncols, 0, sp->d_s, (double*)fac.p, stream, true)) != CPF_OK) break;
    if (ncols > 0) SP_CUDA(cudaMemcpyAsync(h_counts.data(), counts.p, (size_t)ncols * sizeof(int), cudaMemcpyDeviceToHost, stream));
    SP_CUDA(cudaStreamSynchronize(stream));     // flags for the caller; the caller's host arrays may go away
#undef SP_CUDA
  } while (0);
  if (rc != CPF_OK) {
    cpf_spline_destroy(sp);
    return rc;
  }
  for (int64_t c = 0; c < ncols; ++c) col_flags[c] = (uint8_t)(h_counts[c] == np ? 1 : (h_counts[c] > 0 ? 2 : 0));
  *out = sp;
  return CPF_OK;
}

int cpf_column_nan_flags(const double* y, int nx, int64_t ncols, int neg_is_nan, uint8_t* col_flags, int device, void* stream_) {
  if (ncols < 0 || nx < 0) return fail(CPF_EINVAL, "cpf_column_nan_flags: negative size");
  if (ncols == 0) return CPF_OK;
  if (!y || !col_flags) return fail(CPF_EINVAL, "cpf_column_nan_flags: null buffer");
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_column_nan_flags: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  ScratchBuf counts;
  std::vector<int> h_counts((size_t)ncols, 0);
  if (nx > 0) {
    CPF_CUDA(counts.alloc((size_t)ncols * sizeof(int), stream));
    CPF_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)ncols * sizeof(int), stream));
    const dim3 grid((unsigned)((ncols + 127) / 128), (unsigned)((nx + PADLOG_ROWS - 1) / PADLOG_ROWS));
    nan_count_kernel<<<grid, 128, 0, stream>>>(y, nx, ncols, neg_is_nan ? 1 : 0, (int*)counts.p);
    CPF_CUDA(cudaGetLastError());
    CPF_CUDA(cudaMemcpyAsync(h_counts.data(), counts.p, (size_t)ncols * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  for (int64_t c = 0; c < ncols; ++c) col_flags[c] = (uint8_t)(h_counts[c] == nx ? 1 : (h_counts[c] > 0 ? 2 : 0));
  return CPF_OK;
}

static int spline_eval_impl(const cpf_spline* sp, const double* xq, int nq, int nu, double* out, int on_device, int transposed,
                            void* stream_) {
  if (!sp) return fail(CPF_EINVAL, "cpf_spline_eval: null spline");
  if (nq < 0) return fail(CPF_EINVAL, "cpf_spline_eval: negative query count");
  sp->last_stream = (cudaStream_t)stream_;
  if (nu < 0 || nu > 3) return fail(CPF_EINVAL, "cpf_spline_eval: derivative order %d not in 0..3", nu);
  if (nq == 0 || sp->ncols == 0) return CPF_OK;
  if (!xq || !out) return fail(CPF_EINVAL, "cpf_spline_eval: null buffer");
  if (!transposed && nq > 65535) return fail(CPF_EUNSUPPORTED, "cpf_spline_eval: more than 65535 query points in one call");
  if (transposed && (nq + 31) / 32 > 65535) return fail(CPF_EUNSUPPORTED, "cpf_spline_eval_t: more than 2 M query points in one call");   // (grid.y has room for four times that)
  DeviceGuard guard(sp->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t cells = (size_t)nq * (size_t)sp->ncols;
  ScratchBuf dq, dout;
  const double* d_xq = xq;
  double* d_out = out;
  if (!on_device) {
    CPF_CUDA(dq.alloc(nq * sizeof(double), stream));
    CPF_CUDA(dout.alloc(cells * sizeof(double), stream));
    CPF_CUDA(cudaMemcpyAsync(dq.p, xq, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    d_xq = (const double*)dq.p;
    d_out = (double*)dout.p;
  }
  ScratchBuf qx, qi;
  CPF_CUDA(qx.alloc(2 * (size_t)nq * sizeof(double), stream));
  CPF_CUDA(qi.alloc(nq * sizeof(int), stream));
  spline_query_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(sp->d_x, sp->nx, d_xq, nq, sp->extrap, sp->log_x, sp->xmin_raw, sp->xmax_raw,
                                                          (double*)qx.p, (int*)qi.p);
  if (transposed) {
    dim3 grid((unsigned)((sp->ncols + 31) / 32), (unsigned)((nq + SPLINE_EVALT_Q - 1) / SPLINE_EVALT_Q));
    spline_eval_t_kernel<<<grid, 256, 0, stream>>>(sp->d_x, sp->d_y, sp->d_s, sp->ncols, (const double*)qx.p, (const int*)qi.p, nq, nu,
                                                   sp->log_y, d_out);
  } else {
    dim3 grid((unsigned)((sp->ncols + 127) / 128), (unsigned)((nq + SPLINE_EVAL_Q - 1) / SPLINE_EVAL_Q));
    spline_eval_kernel<<<grid, 128, 0, stream>>>(sp->d_x, sp->d_y, sp->d_s, sp->ncols, (const double*)qx.p, (const int*)qi.p, nq, nu,
                                                 sp->log_y, d_out);
  }
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(out, d_out, cells * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  return CPF_OK;
}

int cpf_spline_eval(const cpf_spline* sp, const double* xq, int nq, int nu, double* out, int on_device, void* stream_) {
  return spline_eval_impl(sp, xq, nq, nu, out, on_device, 0, stream_);
}

int cpf_spline_eval_t(const cpf_spline* sp, const double* xq, int nq, int nu, double* out, int on_device, void* stream_) {
  return spline_eval_impl(sp, xq, nq, nu, out, on_device, 1, stream_);
}

int cpf_spline_eval_rows(const double* x, const double* y, int nx, int64_t rows, const double* xq, int nq, int bc, int window,
                         int extrap, double* out, int on_device, int device, void* stream_) {
  if (nx < 2) return fail(CPF_EINVAL, "cpf_spline_eval_rows: need at least 2 knots, got %d", nx);
  if (rows < 0 || nq < 0) return fail(CPF_EINVAL, "cpf_spline_eval_rows: negative size");
  if (bc != 0 && bc != 1) return fail(CPF_EINVAL, "cpf_spline_eval_rows: bc must be 0 (natural) or 1 (clamped)");
  if (window < 0) return fail(CPF_EINVAL, "cpf_spline_eval_rows: negative window");
  if (rows == 0 || nq == 0) return CPF_OK;
  if (!x || !y || !xq || !out) return fail(CPF_EINVAL, "cpf_spline_eval_rows: null buffer");
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_spline_eval_rows: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int W = (window == 0 || window > nx) ? nx : window;   // the host default is 64 (cosmoprimo_b200/interp.py)
  const int LW = (2 * W + 2 < nx) ? 2 * W + 2 : nx;
  const size_t ycells = (size_t)rows * nx, ocells = (size_t)nq * rows;
  ScratchBuf dx, dy, dq, dout, dw, dwork, dmeta;
  const double *p_x = x, *p_y = y, *p_q = xq;
  double* p_out = out;
  if (!on_device) {
    CPF_CUDA(dx.alloc(nx * sizeof(double), stream));
    CPF_CUDA(dy.alloc(ycells * sizeof(double), stream));
    CPF_CUDA(dq.alloc(nq * sizeof(double), stream));
    CPF_CUDA(dout.alloc(ocells * sizeof(double), stream));
    CPF_CUDA(cudaMemcpyAsync(dx.p, x, nx * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(dy.p, y, ycells * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(dq.p, xq, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    p_x = (const double*)dx.p; p_y = (const double*)dy.p; p_q = (const double*)dq.p; p_out = (double*)dout.p;
  }
  CPF_CUDA(dw.alloc((size_t)nq * LW * sizeof(double), stream));
  CPF_CUDA(dwork.alloc((size_t)nq * 2 * LW * sizeof(double), stream));
  CPF_CUDA(dmeta.alloc((size_t)nq * 3 * sizeof(int), stream));
  const size_t wsmem = 2 * (size_t)LW * sizeof(double);
  const int use_smem = wsmem <= 48 * 1024;
  spline_row_weights_kernel<<<nq, 32, use_smem ? wsmem : 0, stream>>>(p_x, nx, bc, W, LW, p_q, nq, (extrap & 1) ? 1 : 0, use_smem,
                                                                     (double*)dw.p, (double*)dwork.p, (int*)dmeta.p);
  CPF_CUDA(cudaGetLastError());
  const int wpb = 8;
  spline_rows_dot_kernel<<<(unsigned)((rows + wpb - 1) / wpb), 32 * wpb, 0, stream>>>(p_y, nx, rows, (const double*)dw.p,
                                                                                    (const int*)dmeta.p, nq, LW, (extrap & 2) ? 1 : 0, p_out);
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(out, p_out, ocells * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  return CPF_OK;
}

}  // extern "C"
