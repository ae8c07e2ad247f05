/*
 * cpfftlog.h — C ABI of libcpfftlog.so, the B200 (sm_100a) engine for cosmoprimo's FFTLog hot path.
 *
 * Every entry point names the reference interface it replaces (paths relative to cosmodesi/cosmoprimo,
 * `cosmoprimo/...`).  The reference is pure Python and has no FFI for this path; the binding a maintainer would
 * add is the ctypes stub shown in INTEGRATION.md (it is what cosmoprimo_b200/_lib.py does).
 *
 * Conventions
 *   - plain pointers and sizes only; all floating-point data is IEEE fp64, row-major, densely packed;
 *   - a pointer is a HOST pointer unless the matching `*_on_device` flag is non-zero, in which case it is a
 *     device pointer on the plan's device (numpy vs DLPack/__cuda_array_interface__ callers);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Host-pointer calls return after the
 *     result is in the caller's buffer; device-pointer calls are asynchronous on `stream`;
 *   - return value: 0 on success, otherwise one of CPF_E*; cpf_last_error() gives the message (thread-local);
 *   - plans are immutable after creation and may be shared between threads; the caller owns all in/out buffers.
 */
#ifndef CPFFTLOG_H
#define CPFFTLOG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPF_VERSION 100

enum cpf_status {
  CPF_OK = 0,
  CPF_EINVAL = 1,        /* bad argument / shape  -> Python ValueError */
  CPF_ECUDA = 2,         /* CUDA runtime failure  -> Python RuntimeError */
  CPF_ENOMEM = 3,
  CPF_EUNSUPPORTED = 4   /* size outside what the kernels cover -> Python NotImplementedError */
};

/* extrapolation modes of cosmoprimo/fftlog.py:466-505 (`pad`): constant fill, 'edge', 'log' */
enum cpf_extrap { CPF_EXTRAP_CONST = 0, CPF_EXTRAP_EDGE = 1, CPF_EXTRAP_LOG = 2 };

typedef struct cpf_plan cpf_plan;

/* ---- library ------------------------------------------------------------------------------------------------ */
int cpf_version(void);
const char* cpf_last_error(void);
/* number of visible CUDA devices; *count = 0 with CPF_ECUDA when no driver/GPU is present */
int cpf_device_count(int* count);
/* Scratch memory (staging buffers of host-pointer calls, work arrays of cpf_wallish2018 / cpf_spline_eval_rows) comes from a private
 * stream-ordered pool per device that keeps at most CPF_SCRATCH_KEEP_MB (environment, default 2048) MiB between calls; the device's
 * default pool and its attributes are never touched.  cpf_trim synchronises the device and returns all unused scratch to the driver. */
int cpf_trim(int device);
/* process-wide diagnostic counters (tests): launches that used the ticket-counter scheduling / launches that took the static split because
 * their ticket slot was still owned by a launch in flight; -1 for an unknown id */
#define CPF_COUNTER_DYNAMIC_LAUNCHES 0
#define CPF_COUNTER_TICKET_FALLBACKS 1
int64_t cpf_counter(int which);

/* ---- FFTLog plan: replaces the tables FFTlog._setup builds and the engine get_fft_engine returns -------------
 * (fftlog.py:144-184, 119-132, 641-663).  The host computes the tables exactly as the reference does (scipy
 * loggamma) and hands them over once; the plan keeps device copies plus everything derived from them
 * (Hermitian-extended kernel spectrum, twiddles).
 *   n, N            : unpadded / padded sizes (fftlog.py:149-150), N a power of two, 2 <= N <= CPF_MAX_N
 *   P               : nparallel (fftlog.py:134-137)
 *   in_left,out_left: padded_size_in_left / padded_size_out_left (fftlog.py:152-153)
 *   pre     [P*N]        padded_prefactor
 *   u_ri    [P*(N/2+1)*2] padded_u, interleaved (re,im)
 *   post_re [P*N]        padded_postfactor (real part)
 *   post_im [P*N] or NULL imaginary part when the post-factor is complex (`complex=True`, fftlog.py:322-330)
 */
#define CPF_MAX_N 8192
int cpf_plan_create(cpf_plan** plan, int n, int N, int P, int in_left, int out_left,
                    const double* pre, const double* u_ri, const double* post_re, const double* post_im,
                    int device);
int cpf_plan_destroy(cpf_plan* plan);
/* introspection used by the tests: which kernel family a call with these options would run
 * (0 = generic shared-memory radix-2, 1 = register radix-16 fast path) */
int cpf_plan_kernel_family(const cpf_plan* plan, int ex_l_mode, double ex_l_val, int ex_r_mode, double ex_r_val,
                           int keep_padding);

/* ---- FFTLog execute: replaces FFTlog.__call__ (fftlog.py:198-241), i.e. pad -> *pre -> rfft -> *u -> conj ->
 * irfft -> *post -> crop, fused in one launch.
 *   in   : [batch, P, n] if in_has_P else [batch, n] (the same row is fed to all P kernels, fftlog.py:231 broadcast)
 *   out  : [batch, P, n_out] doubles, n_out = keep_padding ? N : n; when the plan has a complex post-factor the
 *          element type is interleaved complex128, i.e. [batch, P, n_out, 2]
 *   ex_*_mode/val : left/right extrapolation (cpf_extrap); val only for CPF_EXTRAP_CONST
 */
int cpf_fftlog(const cpf_plan* plan, const double* in, int64_t batch, int in_has_P,
               int ex_l_mode, double ex_l_val, int ex_r_mode, double ex_r_val, int keep_padding,
               double* out, int in_on_device, int out_on_device, void* stream);

/* ---- unfused engine duck type: replaces NumpyFFTEngine.forward / .backward (fftlog.py:538-544) so that an
 * instance can be handed to the unmodified reference as FFTlog(..., engine=instance) (fftlog.py:663).
 *   cpf_rfft        : in [rows, size] real      -> out [rows, size/2+1, 2]
 *   cpf_irfft_conj  : in [rows, size/2+1, 2]    -> out [rows, size] = irfft(conj(in), n=size)
 */
int cpf_rfft(int size, const double* in, int64_t rows, double* out, int in_on_device, int out_on_device,
             int device, void* stream);
int cpf_irfft_conj(int size, const double* in, int64_t rows, double* out, int in_on_device, int out_on_device,
                   int device, void* stream);

/* ---- batched cubic splines: replaces Interpolator1D.__init__ / __call__ on its numpy path (jax.py:139-196), i.e.
 * scipy.interpolate.CubicSpline(x, fun, axis=0, bc_type='natural') + PPoly evaluation, and the clamped splines of the
 * Wallish2018 filter (bao_filter.py:377-382, 400-402, 420).
 * Column layout as in the reference (axis 0 = knots): y [nx, ncols] row-major, shared strictly increasing x [nx].
 *   bc            : 0 = natural (y''=0 at both ends), 1 = clamped (y'=0 at both ends), 2 = not-a-knot (nx >= 4; the ends of
 *                   FITPACK's interpolating splines, i.e. of RectBivariateSpline(s=0) behind Interpolator2D, jax.py:213-243)
 *   log_x, log_y  : fit in log10(x) / log10(y) and return 10**spline (interp_x='log' / interp_fun='log', jax.py:152-153,189-191)
 *   extrap        : 0 => NaN outside [x[0], x[nx-1]] (jax.py:188-192), 1 => extend the end polynomials
 * The handle owns device copies of the (transformed) knots and the fitted slopes.
 *   cpf_spline_eval : out [nq, ncols] = nu-th derivative (0..3) of the spline at xq [nq]
 */
typedef struct cpf_spline cpf_spline;
int cpf_spline_create(cpf_spline** spline, const double* x, const double* y, int nx, int64_t ncols, int bc,
                      int log_x, int log_y, int extrap, int on_device, int device, void* stream);
/* The construction of PowerSpectrumInterpolator1D / 2D with extrap_pk='log' in ONE pass over the table (replaces `_pad_log`,
 * interpolator.py:42-87, the `10**` of :349-351 and the log10 / NaN screening of Interpolator1D, jax.py:152-172): natural spline of
 * log10(y) in log10(x) on nx + 4 knots, where
 *   x_padded [nx + 4] : HOST, the padded wavenumbers as the reference forms them (two continuation knots, the nx tabulated
 *                       wavenumbers, two continuation knots), positive and strictly increasing;
 *   y [nx, ncols]     : the tabulated spectra (host or device); rows 0, 1, nx + 2, nx + 3 of the fitted table continue log10(y) as the
 *                       straight lines through its two lowest / two highest rows (the reference's power-law extrapolation);
 *   col_flags [ncols] : HOST, out: 1 = every cell of the column is NaN (a negative sample counts as NaN): the column evaluates to NaN;
 *                       2 = some cells are: the reference's fit is poisoned as a whole (jax.py:166-172), the caller decides; 0 = clean.
 * Synchronises the stream (the flags are a host result). */
int cpf_spline_create_padlog(cpf_spline** spline, const double* x_padded, const double* y, int nx, int64_t ncols, int extrap,
                             uint8_t* col_flags, int on_device, int device, void* stream);
/* The NaN screening Interpolator1D applies to a table before fitting it (jax.py:161-172), for DEVICE tables y [nx, ncols]: col_flags
 * [ncols] (HOST, out) as above -- 1: every cell of the column is NaN (or negative, with neg_is_nan = the log10-ordinate rule), 2: some
 * are, 0: none.  One read of the table; synchronises the stream. */
int cpf_column_nan_flags(const double* y, int nx, int64_t ncols, int neg_is_nan, uint8_t* col_flags, int device, void* stream);
int cpf_spline_eval(const cpf_spline* spline, const double* xq, int nq, int nu, double* out, int on_device,
                    void* stream);
/* same values, transposed: out [ncols, nq] -- one row per spline, the layout cpf_fftlog reads (saves the `.T` copy of
 * `TophatVariance(k)(pk(k).T)`, interpolator.py:288, 602, 983) */
int cpf_spline_eval_t(const cpf_spline* spline, const double* xq, int nq, int nu, double* out, int on_device,
                      void* stream);
int cpf_spline_destroy(cpf_spline* spline);

/* Row layout, no handle: natural (bc=0) or clamped (bc=1) cubic splines along the LAST axis of y [rows, nx] (the layout
 * cpf_fftlog writes) on shared knots x [nx], evaluated at xq [nq] -> out [nq, rows].  Replaces
 * `Interpolator1D(s, var.T, assume_sorted=True)(r)` of integrate_sigma_r2(method='fftlog') (interpolator.py:288-289)
 * without the two transposes and without a global fit: the value at xq is a weighted sum of the ordinates whose
 * weights depend on (x, xq) only and decay like 0.27^distance, so the slope system is solved on `window` knots either
 * side of the bracketing interval (truncation ~0.27^window; window = 0 or >= nx: all knots = the full solve).
 *   extrap : bit 0: 0 => NaN outside [x[0], x[nx-1]], 1 => extend the end polynomials; bit 1 (value 2): write the square root of the
 *            value (sigma(r) from the variance, `sigma_r = integrate_sigma_r2(...)**0.5`, interpolator.py:573)
 */
int cpf_spline_eval_rows(const double* x, const double* y, int nx, int64_t rows, const double* xq, int nq, int bc,
                         int window, int extrap, double* out, int on_device, int device, void* stream);

/* ---- DST-II / DST-III (orthonormal) along axis 0: replaces scipy.fftpack.dst(type=2, norm='ortho', axis=0) and
 * idst(type=2, norm='ortho', axis=0) at bao_filter.py:372, 412.  data [nx, ncols] row-major, nx a power of two.
 */
int cpf_dst(int type /*2 or 3*/, const double* in, int nx, int64_t ncols, double* out,
            int on_device, int device, void* stream);

/* ---- Wallish2018 no-wiggle filter: replaces Wallish2018PowerSpectrumBAOFilter._compute (bao_filter.py:361-423)
 * given the two spline evaluations of the input spectrum that the reference makes:
 *   klin [nlin] (= linspace(extrap_kmin, 2, 4096)), pklin [nlin, ncols]    (bao_filter.py:364-369)
 *   kout [nk]   (= self.k),                         pkout [nk, ncols]      (bao_filter.py:90-102)
 *   pknow [nk, ncols]  result (bao_filter.py:423)
 *   boxes [ncols, 4] (optional, may be NULL): the even/odd cut boxes (ibox_even, ibox_odd, bao_filter.py:394-395)
 */
int cpf_wallish2018(const double* klin, const double* pklin, int nlin, const double* kout, const double* pkout,
                    int nk, int64_t ncols, double* pknow, int32_t* boxes,
                    int on_device, int device, void* stream);
/* The same filter with the linear-grid spectra handed over one ROW per spectrum, pklin_rows [ncols, nlin] -- the layout cpf_spline_eval_t
 * writes (`pk_interpolator(klin)` evaluated as rows): the kernel fetches a pair of spectra with two 32 KB bulk copies instead of 4096
 * strided 16-byte requests.  pkout / pknow / boxes as above (reference layout); results are bit-identical to cpf_wallish2018.
 * klin and kout are HOST arrays here whatever on_device says (the knot selection needs them on the host; everything derived from them is
 * cached per grid pair), so with device spectra the call neither copies nor synchronises: it is asynchronous on the stream.
 * pklin_rows must be 16-byte aligned. */
int cpf_wallish2018_rows(const double* klin, const double* pklin_rows, int nlin, const double* kout, const double* pkout,
                         int nk, int64_t ncols, double* pknow, int32_t* boxes, int on_device, int device, void* stream);

/* ---- on-device Eisenstein & Hu linear P(k, z): replaces, for B flat LCDM cosmologies without massive neutrinos,
 * `Cosmology(..., engine='eisenstein_hu').get_fourier().pk_interpolator()(k, z)` (eisenstein_hu.py:34-92 coefficients,
 * 241-283 transfer function, 189-214 primordial spectrum, 321-324 P(k), 115-153 growth factor (znorm=0) / growth rate on
 * the background of cosmology.py:1675-1760) and writes rows in the layout cpf_fftlog reads.
 *   params  [B, 5]  (h, omega_b, omega_cdm, n_s, A_s) per cosmology
 *   z       [B, nz] redshifts of every cosmology, 1 <= nz <= 1024 (the transfer function is evaluated once per
 *           cosmology); NULL with nz = 1: redshift 0
 *   k       [nk]    wavenumbers, h/Mpc
 *   T_cmb, omega_r (= Omega0_r h^2: photons + massless neutrinos, cosmology.py:355-367), k_pivot [1/Mpc]
 *   kaiser  0: out [B, nz, nk] = P(k, z);  1: out [B, nz, 3, nk] = Kaiser multipoles ell = 0, 2, 4 with f = growth_rate(z)
 *   derived [B, nz, 4] or NULL: rs_drag [Mpc/h], z_drag, growth_factor(z, znorm=0)^2, growth_rate(z)
 */
int cpf_eh_pk(const double* params, const double* z, int64_t B, int nz, const double* k, int nk, double T_cmb, double omega_r,
              double k_pivot, int kaiser, double* out, double* derived, int on_device, int device, void* stream);

/* ---- measurement helper: peak fp64 FMA rate of the device (DFMA chains), in FLOP/s; used by bench.py for the
 * fp64 roofline denominator that MEASURED_PEAKS.json lacks. */
int cpf_measure_fp64_peak(int device, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CPFFTLOG_H */
