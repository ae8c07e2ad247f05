"""2-D (k, z) interpolators: golden vectors from the reference's PowerSpectrumInterpolator2D /
CorrelationFunctionInterpolator2D / Interpolator2D (tools/make_golden.py::make_interp2d)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import spline_oracle as SO


def golden():
    return np.load(os.path.join(GOLDEN_DIR, 'interp2d_golden.npz'))


def close_with_nans(out, ref, rtol, atol=0.):
    out, ref = np.asarray(out), np.asarray(ref)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    m = np.isfinite(ref)
    np.testing.assert_allclose(out[m], ref[m], rtol=rtol, atol=atol)


def test_oracle_2d_and_factorisation_vs_reference():
    d = golden()
    i2 = SO.interpolator2d(d['k'], d['z'], d['pk'], interp_x='log', interp_fun='log')
    close_with_nans(i2(d['kq'], d['zq']), d['i2_grid'], rtol=1e-13)
    close_with_nans(i2(d['kp'], d['zp'], grid=False), d['i2_pairs'], rtol=1e-13)
    # not-a-knot factorisation == FITPACK's bicubic interpolant (inside the table)
    inside_k, inside_z = (d['kq'] >= d['k'][0]) & (d['kq'] <= d['k'][-1]), (d['zq'] >= 0.) & (d['zq'] <= 3.)
    kq, zq = d['kq'][inside_k], d['zq'][inside_z]
    fact = 10**SO.interpolator2d_factorised(np.log10(d['k']), d['z'], np.log10(d['pk']), np.log10(kq), zq)
    np.testing.assert_allclose(fact, d['i2_grid'][np.ix_(inside_k, inside_z)], rtol=1e-11)


@pytest.mark.gpu
def test_interpolator2d_golden():
    torch = pytest.importorskip('torch')
    from cosmoprimo_b200.interp import Interpolator2D
    d = golden()
    i2 = Interpolator2D(d['k'], d['z'], d['pk'], interp_x='log', interp_fun='log')
    close_with_nans(i2(d['kq'], d['zq']), d['i2_grid'], rtol=1e-10)
    close_with_nans(i2(d['kp'], d['zp'], grid=False), d['i2_pairs'], rtol=1e-10)
    lin = Interpolator2D(np.log(d['k']), d['z'], np.log(d['pk']), extrap=True)
    got, ref = lin(np.log(d['kq']), d['zq']), d['i2_lin_extrap']
    assert np.max(np.abs(got - ref)) < 1e-10 * np.max(np.abs(ref))         # outside the table: FITPACK clamps to the edge
    dev = Interpolator2D(d['k'], d['z'], torch.from_numpy(d['pk']).cuda(), interp_x='log', interp_fun='log')
    out = dev(d['kq'], d['zq'])
    assert isinstance(out, torch.Tensor) and out.is_cuda
    close_with_nans(out.cpu().numpy(), d['i2_grid'], rtol=1e-10)
    # unsorted input, shapes, dtype
    perm_k, perm_z = np.random.default_rng(0).permutation(d['k'].size), np.random.default_rng(1).permutation(d['z'].size)
    shuffled = Interpolator2D(d['k'][perm_k], d['z'][perm_z], d['pk'][np.ix_(perm_k, perm_z)], interp_x='log', interp_fun='log')
    close_with_nans(shuffled(d['kq'], d['zq']), d['i2_grid'], rtol=1e-10)
    assert i2(0.1, 0.5).shape == () and i2(np.ones((2, 3)), np.ones(4)).shape == (2, 3, 4)
    assert i2(np.ones(3, dtype='f4'), np.ones(2, dtype='f4')).dtype == np.float32


@pytest.mark.gpu
def test_power_spectrum_interpolator2d_golden():
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator2D, CorrelationFunctionInterpolator2D, PowerSpectrumInterpolator1D
    d = golden()
    interp = PowerSpectrumInterpolator2D(d['k'], d['z'], d['pk'])
    close_with_nans(interp(d['kq'], d['zq']), d['p2_grid'], rtol=1e-10)
    close_with_nans(interp(d['kp'], d['zp'], grid=False), d['p2_pairs'], rtol=1e-10)
    np.testing.assert_allclose(interp.sigma_rz(d['r'], d['zs']), d['p2_sigma_rz'], rtol=1e-10)
    np.testing.assert_allclose(interp.sigma8_z(d['zs']), d['p2_sigma8_z'], rtol=1e-10)
    # finite differences with dz = 1e-3 amplify the 1e-13 agreement of sigma by 1/dz
    np.testing.assert_allclose(interp.growth_rate_rz(d['r'], d['zs']), d['p2_growth_rate_rz'], rtol=1e-8, atol=1e-9)
    one = interp.to_1d(0.55)
    assert isinstance(one, PowerSpectrumInterpolator1D)
    close_with_nans(one(d['kq']), d['p2_to_1d'], rtol=1e-10)
    xi = interp.to_xi()
    assert isinstance(xi, CorrelationFunctionInterpolator2D)
    np.testing.assert_allclose(xi.s, d['xi_s'], rtol=1e-13)
    got = xi(d['sq'], d['zq'][:5])
    assert np.max(np.abs(got - d['p2_xi'])) < 1e-10 * np.max(np.abs(d['p2_xi']))
    back = xi.to_pk(extrap_pk='lin')(np.geomspace(1e-3, 10., 30), d['zs'])
    np.testing.assert_allclose(back, d['p2_xi_back'], rtol=1e-8)
    interp.rescale_sigma8(0.8)
    np.testing.assert_allclose(interp(d['kq'][:20], d['zs']), d['p2_rescaled'], rtol=1e-10)
    assert abs(float(interp.sigma8_z(0.)) - 0.8) < 1e-12
    # single column + growth_factor_sq callable
    z, D2 = d['z'], None
    from cosmoprimo_b200 import synthetic as S
    D2 = S.growth_factor(z, 0.3137721026737606, 0.6736)**2
    gf = lambda zz: np.interp(zz, z, D2 / D2[0])
    interp = PowerSpectrumInterpolator2D(d['k'], 0., S.eh_pk(d['k']), growth_factor_sq=gf)
    close_with_nans(interp(d['kq'], d['zq']), d['g_grid'], rtol=1e-10)
    np.testing.assert_allclose(interp.sigma_rz(d['r'], d['zs']), d['g_sigma_rz'], rtol=1e-10)
    np.testing.assert_allclose(interp.growth_rate_rz(d['r'], d['zs']), d['g_growth_rate_rz'], rtol=1e-8, atol=1e-9)   # exactly 0 at z = 0 in the reference (flat growth callable below the table)
    got = interp.to_xi()(d['sq'], d['zs'])
    assert np.max(np.abs(got - d['g_xi'])) < 1e-10 * np.max(np.abs(d['g_xi']))


@pytest.mark.gpu
def test_wallish2018_on_2d_interpolator():
    """BAO filter fed by a 2-D (k, z) interpolator (ref bao_filter.py:92-102, 115-145): one column per redshift of the table,
    no growth factor; smooth interpolators come back 2-D.  Checked against the oracle on the arrays our interpolator returns
    (the filter's argmax boxes need bit-identical inputs)."""
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator2D, CorrelationFunctionInterpolator2D
    from cosmoprimo_b200.bao_filter import PowerSpectrumBAOFilter
    from oracle import wallish_oracle as WO
    d = golden()
    interp = PowerSpectrumInterpolator2D(d['k'], d['z'], d['pk'])
    filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
    assert filt.pknow.shape == (1024, d['z'].size) and filt.pk.shape == filt.pknow.shape
    klin = np.linspace(interp.extrap_kmin, 2., 4096)
    ref, dbg = WO.wallish2018(klin, interp(klin, interp.z, ignore_growth=True), filt.k, interp(filt.k, interp.z, ignore_growth=True), return_debug=True)
    assert np.array_equal(filt._boxes, dbg['boxes'])
    assert np.max(np.abs(filt.pknow / ref - 1.)) < 1e-10
    smooth = filt.smooth_pk_interpolator()
    assert isinstance(smooth, PowerSpectrumInterpolator2D)
    inner = slice(10, -10)
    np.testing.assert_allclose(smooth(filt.k[inner], d['z']), filt.pknow[inner], rtol=1e-9)
    xi = filt.smooth_xi_interpolator()
    assert isinstance(xi, CorrelationFunctionInterpolator2D)
    sq = np.geomspace(1., 150., 20)
    assert np.isfinite(xi(sq, d['z'])).all()
    # no BAO peak left: the smooth correlation function at z = 0 has no local maximum between 80 and 130 Mpc/h, the input does
    sfine = np.linspace(80., 130., 200)
    peak = lambda v: np.any((v[1:-1] > v[:-2]) & (v[1:-1] > v[2:]))
    z0 = np.zeros(1)
    assert peak(interp.to_xi()(sfine, z0)[:, 0] * sfine**2) and not peak(xi(sfine, z0)[:, 0] * sfine**2)


@pytest.mark.gpu
def test_sigma_r_against_quadrature():
    """Known answer independent of the reference's FFTLog: sigma(r) by adaptive quadrature of the top-hat integral (the reference's
    own check, tests/test_fftlog.py:16-23, 134-146, rtol 1e-5), for the FFTLog + row-spline path and for the interpolator method."""
    from scipy import integrate
    from cosmoprimo_b200 import synthetic as S
    from cosmoprimo_b200.fftlog import TophatVariance
    from cosmoprimo_b200.interp import spline_eval_rows
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D
    ktab = np.geomspace(1e-6, 1e2, 1500)
    interp = PowerSpectrumInterpolator1D(ktab, S.eh_pk(ktab))

    def wtophat(x):
        return 3. * (np.sin(x) - x * np.cos(x)) / x**3

    def sigma_quad(r, kmin=1e-6, kmax=100., epsrel=1e-5):
        integrand = lambda logk: float(interp(np.exp(logk))) * (wtophat(r * np.exp(logk)) * np.exp(logk))**2 * np.exp(logk)
        return np.sqrt(1. / 2. / np.pi**2 * integrate.quad(integrand, np.log(kmin), np.log(kmax), epsrel=epsrel, limit=400)[0])

    r = np.linspace(1., 20., 10)
    ref = np.array([sigma_quad(rr) for rr in r])
    k = np.logspace(-5, 2, 1000)
    r2, var = TophatVariance(k, lowring=True)(interp(k))
    np.testing.assert_allclose(np.sqrt(spline_eval_rows(r2, var[None, :], r)[:, 0]), ref, rtol=1e-5)
    np.testing.assert_allclose(interp.sigma_r(r), ref, rtol=1e-5)


@pytest.mark.gpu
def test_nan_tables_and_bounds_contract():
    """The reference's test_nan (tests/test_interpolator.py:328-337) and the bounds part of test_extrap_2d (:232-300)."""
    from cosmoprimo_b200 import synthetic as S
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D, PowerSpectrumInterpolator2D
    k = np.logspace(-4, 2, 1000)
    pk = k**2
    pk[:2] *= -1                                               # log10 of a negative number: the whole fit is poisoned
    with np.errstate(invalid='ignore'):
        assert np.isnan(PowerSpectrumInterpolator1D(k, pk)(k)).all()
        z = np.linspace(0., 2., 4)
        assert np.isnan(PowerSpectrumInterpolator2D(k, z, pk[..., None][..., [0] * len(z)])(k, z=1.)).all()
    # bounds: NaN outside the extrapolation range / redshift table, ValueError with bounds_error
    z = np.linspace(0., 4., 10)
    D2 = S.growth_factor(z, 0.3137721026737606, 0.6736)**2
    tab = S.eh_pk(k)[:, None] * D2
    k_extrap = np.logspace(-6, 3, 1000)
    k_eval = k_extrap[1:-1]
    interp = PowerSpectrumInterpolator2D(k, z, tab, extrap_kmin=k_extrap[0], extrap_kmax=k_extrap[-1])
    assert np.isfinite(interp(k_eval, z)).all()
    assert np.isnan(interp(k_eval[0] / 2., z)).all() and np.isnan(interp(k_eval[-1] * 2., z)).all()
    assert np.isnan(interp(k_eval, z[-1] * 2.)).all()
    for kk, zz in [(k_eval / 2., z), (k_eval * 2., z), (k_eval, z * 2.)]:
        with pytest.raises(ValueError):
            interp(kk, zz, bounds_error=True)
    xi = interp.to_xi()
    s_eval = xi.s
    assert np.isfinite(xi(s_eval, z)).all()
    assert np.isnan(xi(s_eval[0] / 2., z)).all() and np.isnan(xi(s_eval[-1] * 2., z)).all() and np.isnan(xi(s_eval, z[-1] * 2.)).all()
    for ss, zz in [(s_eval / 2., z), (s_eval * 2., z), (s_eval, z * 2.)]:
        with pytest.raises(ValueError):
            xi(ss, zz, bounds_error=True)
    # round trip xi -> P(k) at z = 0 within 1 % over the tabulated range (tests/test_interpolator.py:300)
    back = xi.to_pk()
    sel = (k > 1e-3) & (k < 10.)
    np.testing.assert_allclose(back(k[sel], 0.), tab[sel, 0], rtol=1e-2)
