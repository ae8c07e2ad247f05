"""GPU parity of the spline / interpolator / DST / Wallish2018 entry points (through the ctypes C ABI) against vectors
produced by the unmodified reference and against the oracles on seeded inputs."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden
from cosmoprimo_b200 import _lib, synthetic as S
from cosmoprimo_b200.interp import Interpolator1D, spline_eval_rows
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D, CorrelationFunctionInterpolator1D
from cosmoprimo_b200.bao_filter import PowerSpectrumBAOFilter, Wallish2018PowerSpectrumBAOFilter
from oracle import spline_oracle as SO
from oracle import wallish_oracle as WO

pytestmark = pytest.mark.gpu

SPLINE_CASES = list(range(len(load_golden('spline_golden.npz').cases)))


def close_with_nans(out, ref, rtol, atol=0.):
    out, ref = np.asarray(out), np.asarray(ref)
    assert out.shape == ref.shape and out.dtype == ref.dtype, (out.shape, ref.shape, out.dtype, ref.dtype)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    m = np.isfinite(ref)
    assert np.array_equal(np.isinf(out), np.isinf(ref))
    np.testing.assert_allclose(out[m], ref[m], rtol=rtol, atol=atol)


@pytest.mark.parametrize('idx', SPLINE_CASES)
def test_spline_golden(idx):
    g = load_golden('spline_golden.npz')
    case, d = g.cases[idx], g.data
    ref = d['s{}'.format(idx)]
    if case['kind'] == 'interp1d':
        interp = Interpolator1D(d[case['x']], d[case['y']], interp_x=case['interp_x'], interp_fun=case['interp_fun'], extrap=case['extrap'],
                                assume_sorted=case.get('assume_sorted', False))
        out = interp(d[case['xq']], dx=case['dx'])
        # values are O(1..1e3) with knot spacing down to 1e-3: derivatives amplify rounding by 1/dx
        close_with_nans(out, ref, rtol=1e-10 if case['dx'] == 0 else 1e-7, atol=1e-12 * np.nanmax(np.abs(ref[np.isfinite(ref)])))
    else:
        interp = PowerSpectrumInterpolator1D(d[case['k']], d[case['pk']])
        close_with_nans(interp(d[case['keval']]), ref, rtol=1e-10)
        r = d['r']
        close_with_nans(interp.sigma_r(r), d['s{}_sigma_r'.format(idx)], rtol=1e-10)
        close_with_nans(np.asarray(interp.sigma8()), d['s{}_sigma8'.format(idx)], rtol=1e-10)
        xi = interp.to_xi()
        assert isinstance(xi, CorrelationFunctionInterpolator1D)
        np.testing.assert_allclose(xi.s, d['s{}_xi_s'.format(idx)], rtol=1e-13)
        xi_ref = d['s{}_xi'.format(idx)]
        got = xi(d['s{}_seval'.format(idx)])
        assert got.shape == xi_ref.shape
        assert np.max(np.abs(got - xi_ref)) < 1e-10 * np.max(np.abs(xi_ref))
        if 's{}_pk_back'.format(idx) in d.files:
            back = xi.to_pk()(np.geomspace(1e-2, 10., 30))
            np.testing.assert_allclose(back, d['s{}_pk_back'.format(idx)], rtol=1e-8)


def test_spline_clamped_and_derivatives_vs_oracle():
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0., 50., 777))
    y = np.cos(x)[:, None] * rng.uniform(0.5, 2., (777, 33))
    xq = rng.uniform(x[0], x[-1], 500)
    for bc in ['natural', 'clamped']:
        s = SO.cubic_spline_slopes(x, y, bc)
        interp = Interpolator1D(x, y, bc_type=bc, assume_sorted=True)
        for nu in [0, 1, 2, 3]:
            ref = SO.cubic_spline_eval(x, y, s, xq, nu=nu)
            out = interp(xq, dx=nu)
            assert np.max(np.abs(out - ref)) < 1e-9 * np.max(np.abs(ref)), (bc, nu)


def test_padlog_matches_the_composed_construction():
    """`Interpolator1D.padlog` (one pass: logarithms, continuation knots, NaN screening, fit) against the composition it replaces --
    `_pad_log`, `10**`, `Interpolator1D(interp_x='log', interp_fun='log')` (ref interpolator.py:42-87, 343-351; jax.py:152-172) -- on host
    and device tables: clean columns, an all-NaN column (passes through as NaN), a negative sample (poisons the whole fit), a zero."""
    torch = pytest.importorskip('torch')
    from cosmoprimo_b200.interpolator import _pad_log, _pad_log_knots
    rng = np.random.default_rng(5)
    k = np.geomspace(2e-5, 40., 257)
    pk = S.eh_pk(k, S.lhs_cosmologies(9, seed=3)).T.copy()                    # (nk, 9)
    kq = np.concatenate([[1e-7, 1e2], np.geomspace(1e-7, 1e2, 301), [5e-8, 2e2]])

    def composed(table):
        kk, pp = _pad_log(k, table)
        return Interpolator1D(10**kk, 10**pp, interp_x='log', interp_fun='log', assume_sorted=True)

    def fused(table):
        logk, lo, hi = _pad_log_knots(k)
        return Interpolator1D.padlog(10**np.concatenate([lo, logk, hi]), table)

    ref = composed(pk)(kq)
    for table in (pk, torch.from_numpy(pk).cuda()):
        out = fused(table)(kq)
        out = out.cpu().numpy() if hasattr(out, 'cpu') else out
        close_with_nans(out, ref, rtol=1e-12)
    assert np.isnan(ref[-2:]).all() and np.isfinite(ref[:2]).all()             # range ends are inside, beyond them NaN
    # all-NaN column: that column NaN, the others untouched
    t = pk.copy(); t[:, 4] = np.nan
    with np.errstate(invalid='ignore'):
        close_with_nans(fused(torch.from_numpy(t).cuda())(kq).cpu().numpy(), composed(t)(kq), rtol=1e-12)
        assert np.isnan(fused(t)(kq)[:, 4]).all() and np.isfinite(fused(t)(kq)[2:-2, 3]).all()
        # one negative sample: the reference's fit is poisoned as a whole
        t = pk.copy(); t[100, 2] = -1.
        out = fused(torch.from_numpy(t).cuda())(kq).cpu().numpy()
        assert np.isnan(out).all() and np.isnan(composed(t)(kq)).all()
    # shapes: 1-D table, trailing axes, float32 queries
    assert fused(pk[:, 0])(kq).shape == kq.shape and fused(pk.reshape(257, 3, 3))(kq[:5]).shape == (5, 3, 3)
    assert fused(pk)(kq.astype('f4')).dtype == np.float32
    # through the public class: device table in, device result out, same numbers as the host table
    a = PowerSpectrumInterpolator1D(k, pk)(kq[2:-2])
    b = PowerSpectrumInterpolator1D(k, torch.from_numpy(pk).cuda())(kq[2:-2])
    assert isinstance(b, torch.Tensor) and np.array_equal(a, b.cpu().numpy())
    # sigma_r: the square root is taken by the kernel on device rows
    sa = PowerSpectrumInterpolator1D(k, pk).sigma_r(np.array([4., 8., 12.]))
    sb = PowerSpectrumInterpolator1D(k, torch.from_numpy(pk).cuda()).sigma_r(np.array([4., 8., 12.]))
    np.testing.assert_allclose(sb.cpu().numpy(), sa, rtol=1e-13)


def test_spline_device_buffers_and_shapes():
    torch = pytest.importorskip('torch')
    g = load_golden('spline_golden.npz').data
    x, y, xq = g['x'], g['y'], g['xq']
    host = Interpolator1D(x, y)(xq)
    dev = Interpolator1D(x, torch.from_numpy(y).cuda())
    out = dev(xq)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), host, equal_nan=True)
    out = dev(torch.from_numpy(xq).cuda().reshape(8, 8))
    assert tuple(out.shape) == (8, 8, 7)
    # transposed evaluation (rows = splines) is the same numbers, on host and device tables, odd sizes
    assert np.array_equal(Interpolator1D(x, y).eval_rows(xq), host.T, equal_nan=True)
    rows = dev.eval_rows(xq[:37])
    assert isinstance(rows, torch.Tensor) and tuple(rows.shape) == (7, 37)
    assert np.array_equal(rows.cpu().numpy(), host[:37].T, equal_nan=True)
    # shape / dtype contract of the reference tests (tests/test_interpolator.py:8-32)
    interp = Interpolator1D(x, y[:, 0])
    assert interp(1.).shape == () and interp(np.array([])).shape == (0,) and interp(np.ones((2, 3))).shape == (2, 3)
    assert interp(np.ones(4, dtype='f4')).dtype == np.float32 and interp(np.ones(4)).dtype == np.float64
    with pytest.raises(ValueError):
        interp(np.array([100.]), bounds_error=True)


def test_spline_eval_rows_vs_oracle():
    """cpf_spline_eval_rows (windowed weights, rows layout) == Interpolator1D(s, var.T)(r) of the reference (interpolator.py:289)."""
    torch = pytest.importorskip('torch')
    n, B = 2048, 37
    k = np.geomspace(1e-5, 1e2, n)
    pk = S.eh_pk(k, S.lhs_cosmologies(B, seed=5))
    from oracle import fftlog_oracle as FO
    s, var = FO.execute(FO.plan_tophat_variance(k), pk)
    r = np.concatenate([np.linspace(1., 20., 10), [s[0], s[-1], 0.5 * (s[0] + s[1]), 8., s[0] * 0.5, s[-1] * 2, np.nan]])
    ref = SO.interpolator1d(s, var.T, assume_sorted=True)(r)
    for window in [0, 128, 64, 40]:
        out = spline_eval_rows(s, var, r, window=window)
        assert out.shape == (r.size, B)
        close_with_nans(out, ref, rtol=1e-11)
    out = spline_eval_rows(s, torch.from_numpy(var).cuda(), r)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    close_with_nans(out.cpu().numpy(), ref, rtol=1e-11)
    # clamped ends, coarse random grid (full solve since nx < 2 * window + 2), extrapolation
    rng = np.random.default_rng(11)
    x = np.sort(rng.uniform(0., 10., 50))
    y = rng.normal(size=(5, 50))
    xq = np.concatenate([rng.uniform(x[0], x[-1], 40), [x[0] - 0.1, x[-1] + 0.2]])
    for bc in ['natural', 'clamped']:
        sl = SO.cubic_spline_slopes(x, y.T, bc)
        ref = SO.cubic_spline_eval(x, y.T, sl, xq, extrapolate=True)
        out = spline_eval_rows(x, y, xq, bc_type=bc, extrap=True)
        assert np.max(np.abs(out - ref)) < 1e-12 * np.max(np.abs(ref)), bc
    assert spline_eval_rows(x, y[:0], xq).shape == (xq.size, 0)


def fake_interpolator(klin, pklin, kout, pkout, extrap_kmin=1e-7, extrap_kmax=1e2):
    """Object with the interpolator duck type that returns the reference's own evaluations (bit-identical inputs)."""
    class Fake(object):
        def __call__(self, k):
            k = np.asarray(k)
            if k.size == klin.size and np.array_equal(k, klin): return pklin
            if k.size == kout.size and np.allclose(k, kout, rtol=1e-14): return pkout
            raise AssertionError('unexpected grid')
    f = Fake()
    f.extrap_kmin, f.extrap_kmax = extrap_kmin, extrap_kmax
    return f


@pytest.mark.parametrize('idx', [0, 1])
def test_wallish_golden(idx):
    d = load_golden('wallish_golden.npz').data
    klin, pklin, kout, pkout = (d['w%d_%s' % (idx, n)] for n in ['klin', 'pklin', 'kout', 'pkout'])
    filt = PowerSpectrumBAOFilter(fake_interpolator(klin, pklin, kout, pkout), engine='wallish2018_cuda')
    assert isinstance(filt, Wallish2018PowerSpectrumBAOFilter)
    np.testing.assert_allclose(filt.k, kout, rtol=1e-14)
    ref, dbg = WO.wallish2018(klin, pklin, kout, pkout, return_debug=True)
    assert np.array_equal(filt._boxes, dbg['boxes'])                   # the four argmax boxes of every column
    assert filt.pknow.shape == d['w%d_pknow' % idx].shape
    assert np.max(np.abs(filt.pknow / d['w%d_pknow' % idx] - 1.)) < 1e-10
    assert np.array_equal(filt.pk, pkout)
    # a B-column run equals B single-column runs on the identical arrays
    for col in [0, pklin.shape[1] - 1]:
        one = PowerSpectrumBAOFilter(fake_interpolator(klin, pklin[:, col], kout, pkout[:, col]), engine='wallish2018_cuda')
        assert one.pknow.shape == (kout.size,)
        # not bit-identical: two columns share one complex FFT, so a column's rounding depends on its partner
        assert np.max(np.abs(one.pknow / filt.pknow[:, col] - 1.)) < 1e-10


def test_wallish_seeded_batch_vs_oracle():
    """End to end from tables: our PowerSpectrumInterpolator1D feeds the filter; 65 (odd) LHS cosmologies."""
    ktab = np.geomspace(1e-5, 1e2, 512)
    pk = S.eh_pk(ktab, S.lhs_cosmologies(65, seed=11)).T
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
    klin = np.linspace(interp.extrap_kmin, 2., 4096)
    ref, dbg = WO.wallish2018(klin, interp(klin), filt.k, interp(filt.k), return_debug=True)
    # count the columns whose four box indices differ from the oracle's on bit-identical inputs: expected 0
    bad = np.nonzero(np.any(filt._boxes != dbg['boxes'], axis=1))[0]
    assert bad.size == 0, 'boxes differ in {} of {} columns: {}'.format(bad.size, len(dbg['boxes']), [(int(c), filt._boxes[c].tolist(), dbg['boxes'][c].tolist()) for c in bad[:8]])
    assert np.max(np.abs(filt.pknow / ref - 1.)) < 1e-10
    smooth = filt.smooth_pk_interpolator()
    assert np.allclose(smooth(filt.k[10:-10]), filt.pknow[10:-10], rtol=1e-9)
    # device-resident inputs (the same evaluated arrays, as torch tensors): same kernels, same bits
    torch = pytest.importorskip('torch')
    pklin, pkout = interp(klin), interp(filt.k)

    class DeviceInterp(object):
        extrap_kmin, extrap_kmax = interp.extrap_kmin, interp.extrap_kmax

        def __call__(self, k):
            return torch.from_numpy(pklin if np.size(k) == klin.size else pkout).cuda()

    filt_d = PowerSpectrumBAOFilter(DeviceInterp(), engine='wallish2018')
    assert isinstance(filt_d.pknow, torch.Tensor) and filt_d.pknow.is_cuda
    assert np.array_equal(filt_d._boxes.cpu().numpy(), filt._boxes)
    assert np.array_equal(filt_d.pknow.cpu().numpy(), filt.pknow)
    # and the fully device-resident chain (tables on the GPU).  Its spline evaluations differ from the host chain's in the last bits
    # (device log10 / exp10 of the padded log-log table), and the boxes are discontinuous in the input, so the
    # oracle is run on the arrays the device chain really evaluated: bit-identical inputs => identical boxes, 1e-10 values.
    interp_d = PowerSpectrumInterpolator1D(ktab, torch.from_numpy(pk).cuda())
    filt_dd = PowerSpectrumBAOFilter(interp_d, engine='wallish2018')
    pklin_d, pkout_d = interp_d(klin).cpu().numpy(), interp_d(filt_dd.k).cpu().numpy()
    assert np.max(np.abs(pklin_d / pklin - 1.)) < 1e-11
    ref_d, dbg_d = WO.wallish2018(klin, pklin_d, filt_dd.k, pkout_d, return_debug=True)
    bad_d = np.nonzero(np.any(filt_dd._boxes.cpu().numpy() != dbg_d['boxes'], axis=1))[0]
    assert bad_d.size == 0, 'device chain: boxes differ in {} columns'.format(bad_d.size)
    assert np.max(np.abs(filt_dd.pknow.cpu().numpy() / ref_d - 1.)) < 1e-10


@pytest.mark.parametrize('ncols', [1, 6, 7])
def test_wallish_rows_entry_is_bit_identical(ncols):
    """cpf_wallish2018_rows (linear-grid spectra one row per spectrum, fetched by bulk copies) == cpf_wallish2018 (reference layout) bit for bit,
    boxes included: host arrays (even / odd column counts) and device arrays; and the filter class takes the rows entry for this package's
    interpolators (1-D on host and device tables, 2-D) with the same result as the reference layout."""
    torch = pytest.importorskip('torch')
    lib = _lib.load()
    d = load_golden('wallish_golden.npz').data
    klin, pklin, kout, pkout = (np.ascontiguousarray(d['w0_%s' % n]) for n in ['klin', 'pklin', 'kout', 'pkout'])
    pklin, pkout = (np.ascontiguousarray(np.tile(a, (1, 2))[:, :ncols]) for a in (pklin, pkout))          # the golden has six columns
    assert pklin.shape == (4096, ncols) and pkout.shape == (kout.size, ncols)
    rows = np.ascontiguousarray(pklin.T)
    out = [np.empty_like(pkout) for _ in range(2)]
    boxes = [np.empty((ncols, 4), dtype='i4') for _ in range(2)]
    _lib.check(lib.cpf_wallish2018(klin.ctypes.data, pklin.ctypes.data, 4096, kout.ctypes.data, pkout.ctypes.data, kout.size, ncols, out[0].ctypes.data, boxes[0].ctypes.data, 0, 0, None))
    _lib.check(lib.cpf_wallish2018_rows(klin.ctypes.data, rows.ctypes.data, 4096, kout.ctypes.data, pkout.ctypes.data, kout.size, ncols, out[1].ctypes.data, boxes[1].ctypes.data, 0, 0, None))
    assert np.array_equal(out[0], out[1], equal_nan=True) and np.array_equal(boxes[0], boxes[1])
    t = {name: torch.from_numpy(a).cuda() for name, a in dict(klin=klin, rows=rows, kout=kout, pkout=pkout).items()}
    dout = torch.empty_like(t['pkout'])
    _lib.check(lib.cpf_wallish2018_rows(klin.ctypes.data, t['rows'].data_ptr(), 4096, kout.ctypes.data, t['pkout'].data_ptr(), kout.size, ncols, dout.data_ptr(), None, 1, 0,        # grids: host
                                        torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(dout.cpu().numpy(), out[0], equal_nan=True)


def test_wallish_filter_class_uses_rows_for_own_interpolators():
    torch = pytest.importorskip('torch')
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator2D
    ktab = np.geomspace(1e-4, 50., 300)
    pk = S.eh_pk(ktab, S.lhs_cosmologies(5, seed=11)).T.copy()                              # (nk, 5)
    for table in (pk, torch.from_numpy(pk).cuda()):
        interp = PowerSpectrumInterpolator1D(ktab, table)
        filt = PowerSpectrumBAOFilter(interp, engine='wallish2018_cuda')
        assert filt._evaluate_rows(filt.k) is not None
        klin = np.linspace(interp.extrap_kmin, 2., 4096)
        pklin, pkout = interp(klin), interp(filt.k)
        to_np = lambda a: a.cpu().numpy() if hasattr(a, 'cpu') else np.asarray(a)
        ref = PowerSpectrumBAOFilter(fake_interpolator(klin, to_np(pklin), filt.k, to_np(pkout)), engine='wallish2018_cuda')       # reference layout
        assert np.array_equal(to_np(filt._boxes), ref._boxes)
        assert np.max(np.abs(to_np(filt.pknow) / ref.pknow - 1.)) < 1e-12              # (eval_t and eval differ by the rounding of their exp10 only)
    z = np.linspace(0., 1., 6)
    interp2 = PowerSpectrumInterpolator2D(ktab, z, pk[:, :1] * (1. + z)[None, :]**-2)
    filt2 = PowerSpectrumBAOFilter(interp2, engine='wallish2018_cuda')
    assert filt2.pknow.shape == (filt2.k.size, z.size) and np.isfinite(filt2.pknow).all()
    klin = np.linspace(interp2.extrap_kmin, 2., 4096)
    ref2 = PowerSpectrumBAOFilter(fake_interpolator(klin, interp2(klin, z), filt2.k, interp2(filt2.k, z)), engine='wallish2018_cuda')
    assert np.array_equal(filt2._boxes, ref2._boxes) and np.max(np.abs(filt2.pknow / ref2.pknow - 1.)) < 1e-12


def test_new_entry_points_reject_bad_input():
    """Error behaviour of the round-2 entry points through the C ABI: bad knots / misaligned rows / null buffers give CPF_EINVAL (ValueError in
    the Python layer), never a crash."""
    torch = pytest.importorskip('torch')
    lib = _lib.load()
    y = np.abs(np.random.default_rng(0).standard_normal((8, 3))) + 1.
    flags = np.zeros(3, dtype='u1')
    handle = ctypes.c_void_p()
    x_bad = np.array([1e-7, 1e-6, 1., 2., 3., 2.5, 5., 6., 7., 8., 9., 1e2])          # 8 + 4 knots, not increasing
    with pytest.raises(ValueError):
        _lib.check(lib.cpf_spline_create_padlog(ctypes.byref(handle), x_bad.ctypes.data, y.ctypes.data, 8, 3, 0, flags.ctypes.data, 0, 0, None))
    x_neg = np.array([-1., 1e-6, 1., 2., 3., 4., 5., 6., 7., 8., 9., 1e2])
    with pytest.raises(ValueError):
        _lib.check(lib.cpf_spline_create_padlog(ctypes.byref(handle), x_neg.ctypes.data, y.ctypes.data, 8, 3, 0, flags.ctypes.data, 0, 0, None))
    with pytest.raises(ValueError):
        _lib.check(lib.cpf_spline_create_padlog(ctypes.byref(handle), x_neg.ctypes.data, None, 8, 3, 0, flags.ctypes.data, 0, 0, None))
    assert not handle.value
    # NaN screening on a device table: all-NaN column -> 1, partly NaN -> 2, clean -> 0; negative counts as NaN only with the log rule
    t = torch.from_numpy(y.copy()).cuda()
    t[:, 0] = float('nan'); t[2, 1] = -1.
    _lib.check(lib.cpf_column_nan_flags(t.data_ptr(), 8, 3, 1, flags.ctypes.data, 0, torch.cuda.current_stream().cuda_stream))
    assert flags.tolist() == [1, 2, 0]
    _lib.check(lib.cpf_column_nan_flags(t.data_ptr(), 8, 3, 0, flags.ctypes.data, 0, torch.cuda.current_stream().cuda_stream))
    assert flags.tolist() == [1, 0, 0]
    # rows entry of the Wallish2018 filter: misaligned rows pointer, non-increasing grid
    d = load_golden('wallish_golden.npz').data
    klin, kout = np.ascontiguousarray(d['w0_klin']), np.ascontiguousarray(d['w0_kout'])
    rows = np.ascontiguousarray(np.concatenate([[0.], d['w0_pklin'][:, 0]]))               # rows[1:] is 8 bytes off a 16-byte boundary
    pkout, out = np.ascontiguousarray(d['w0_pkout'][:, :1]), np.empty((kout.size, 1))
    assert rows[1:].ctypes.data % 16 == 8
    rows_d = torch.from_numpy(rows).cuda()
    pk_d, out_d = torch.from_numpy(pkout).cuda(), torch.from_numpy(out).cuda()
    with pytest.raises(ValueError):
        _lib.check(lib.cpf_wallish2018_rows(klin.ctypes.data, rows_d.data_ptr() + 8, 4096, kout.ctypes.data, pk_d.data_ptr(), kout.size, 1, out_d.data_ptr(), None, 1, 0, None))
    kbad = klin.copy(); kbad[100] = kbad[99]
    with pytest.raises(ValueError):
        _lib.check(lib.cpf_wallish2018_rows(kbad.ctypes.data, rows_d.data_ptr(), 4096, kout.ctypes.data, pk_d.data_ptr(), kout.size, 1, out_d.data_ptr(), None, 1, 0, None))


def test_dst_matches_scipy():
    """cpf_dst == scipy.fftpack.dst / idst (type 2, ortho, axis 0) that the reference calls (bao_filter.py:372, 412)."""
    from scipy import fftpack
    lib = _lib.load()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4096, 7))
    out = np.empty_like(x)
    _lib.check(lib.cpf_dst(2, x.ctypes.data, 4096, 7, out.ctypes.data, 0, 0, None))
    ref = fftpack.dst(x, type=2, axis=0, norm='ortho')
    assert np.max(np.abs(out - ref)) < 1e-14 * np.max(np.abs(ref)) * 10
    back = np.empty_like(x)
    _lib.check(lib.cpf_dst(3, out.ctypes.data, 4096, 7, back.ctypes.data, 0, 0, None))
    assert np.max(np.abs(back - fftpack.idst(ref, type=2, axis=0, norm='ortho'))) < 1e-13
    assert np.max(np.abs(back - x)) < 1e-13
    with pytest.raises(NotImplementedError):
        _lib.check(lib.cpf_dst(2, x.ctypes.data, 2048, 7, out.ctypes.data, 0, 0, None))


def test_wallish_output_outside_spliced_knots_is_nan():
    """with extrap_kmin >= 5e-4 (no left splice) or extrap_kmax in (1.5, 2] (no right splice) the reference's final
    CubicSpline(extrapolate=False) returns NaN at the output wavenumbers outside the spliced knots (ref bao_filter.py:415-423)."""
    ktab = np.geomspace(1e-5, 1e2, 512)
    pk = S.eh_pk(ktab, S.lhs_cosmologies(5, seed=2)).T
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    for kmin, kmax in [(1e-3, 1e2), (1e-7, 1.8), (2e-3, 1.9)]:
        klin = np.linspace(kmin, 2., 4096)
        kout = np.geomspace(kmin, kmax, 1024)
        pklin, pkout = interp(klin), interp(kout)
        filt = PowerSpectrumBAOFilter(fake_interpolator(klin, pklin, kout, pkout, extrap_kmin=kmin, extrap_kmax=kmax), engine='wallish2018_cuda')
        ref = WO.wallish2018(klin, pklin, kout, pkout)
        assert np.isnan(ref).any() and np.array_equal(np.isnan(filt.pknow), np.isnan(ref))
        m = np.isfinite(ref)
        assert np.max(np.abs(filt.pknow[m] / ref[m] - 1.)) < 1e-10


def test_wallish_multi_round_evaluation(monkeypatch):
    """The slopes next to the output wavenumbers travel through a slot array in the shared buffer; with more output wavenumbers than slots the
    evaluation runs in rounds.  CPF_WALLISH_SLOT_CAP shrinks the slot array so that the 1024 default wavenumbers need > 10 rounds: same bits."""
    d = load_golden('wallish_golden.npz').data
    klin, pklin, kout, pkout = (d['w0_%s' % n] for n in ['klin', 'pklin', 'kout', 'pkout'])
    one = PowerSpectrumBAOFilter(fake_interpolator(klin, pklin, kout, pkout), engine='wallish2018_cuda').pknow
    monkeypatch.setenv('CPF_WALLISH_SLOT_CAP', '37')
    many = PowerSpectrumBAOFilter(fake_interpolator(klin, pklin, kout, pkout), engine='wallish2018_cuda').pknow
    assert np.array_equal(one, many)
    assert np.max(np.abs(many / d['w0_pknow'] - 1.)) < 1e-10
