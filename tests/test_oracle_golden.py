"""The oracle (oracle/fftlog_oracle.py) pinned against the reference's own outputs and known answers. CPU only."""
import numpy as np
import pytest

from conftest import load_golden, Golden, scale_aware_error
from oracle import fftlog_oracle as O

PLANNERS = {'HankelTransform': O.plan_hankel, 'PowerToCorrelation': O.plan_power_to_correlation,
            'CorrelationToPower': O.plan_correlation_to_power, 'TophatVariance': O.plan_tophat_variance,
            'GaussianVariance': O.plan_gaussian_variance}

CASES = list(range(len(load_golden().cases)))


def oracle_run(golden, idx):
    case = golden.cases[idx]
    pl = PLANNERS[case['cls']](golden.inp(case['grid']), **Golden.ctor_kwargs(case))
    kw = Golden.call_kwargs(case)
    if case['inv']:
        O.invert_plan(pl)
    return pl, O.execute(pl, golden.fun(idx), **kw)


@pytest.mark.parametrize('idx', CASES)
def test_oracle_matches_reference(fftlog_golden, idx):
    """Every golden case: same grids, same tables, same output as the reference numpy engine."""
    case = fftlog_golden.cases[idx]
    pl, (y, g) = oracle_run(fftlog_golden, idx)
    assert (pl['N'], pl['n'], pl['P']) == (case['N'], case['n'], case['P'])
    y_ref, g_ref = fftlog_golden.get(idx, 'y'), fftlog_golden.get(idx, 'g')
    assert y.shape == y_ref.shape and g.shape == g_ref.shape and g.dtype == g_ref.dtype
    np.testing.assert_allclose(y, y_ref, rtol=1e-14, atol=0)
    if case['tables']:
        for name, key in [('pre', 'padded_prefactor'), ('u', 'padded_u'), ('post', 'padded_postfactor')]:
            np.testing.assert_allclose(pl[key], fftlog_golden.get(idx, name), rtol=1e-13, atol=0)
    # same numpy/scipy => bit-identical; another numpy build may differ in the last bits of the FFT
    post = pl['padded_postfactor']
    if not case['callkw'].get('keep_padding', False) and not case['inv']:
        post = post[..., pl['out_left']:pl['out_left'] + pl['n']]
    if post.shape[-1] != g_ref.shape[-1]:   # inv() with its unpadded tables
        post = np.ones(g_ref.shape[-1])
    assert scale_aware_error(g, g_ref, post) < 1e-13


def test_analytic_hankel_pair():
    """The reference's known-answer test (tests/test_fftlog.py:56-89): 1/(1+x^2)^1.5 <-> exp(-y), nu=0, q=1."""
    x = np.logspace(-3, 3, num=60, endpoint=False)
    f = 1 / (1 + x**2)**1.5
    pl = O.plan_hankel(x, nu=0, q=1, lowring=True)
    y, g = O.execute(pl, f, extrap='log')
    assert np.allclose(g, np.exp(-y), rtol=1e-8, atol=1e-8)
    O.invert_plan(pl)
    x2, f2 = O.execute(pl, g, extrap='log')
    assert np.allclose(f2, f, rtol=1e-7, atol=1e-7)
    y = np.logspace(-4, 2, num=60, endpoint=False)
    pl = O.plan_hankel(y, nu=0, q=1, lowring=True)
    x, f = O.execute(pl, np.exp(-y), extrap='log')
    assert np.allclose(f, 1 / (1 + x**2)**1.5, rtol=1e-10, atol=1e-10)


def test_pad_known_answers():
    """tests/test_fftlog.py:26-53 and SURVEY Appendix A.7."""
    assert np.allclose(O.pad_last([1., 2., 4., 8.], 2, 3, 'log'), [.25, .5, 1, 2, 4, 8, 16, 32, 64])
    assert np.allclose(O.pad_last([1., 2., 4., 8.], 2, 3, (0, 'log')), [0, 0, 1, 2, 4, 8, 16, 32, 64])
    assert np.allclose(O.pad_last(np.ones((6, 6)), 4, 3, 'edge'), np.ones((6, 13)))
    x = np.logspace(-3, 3, num=7, endpoint=True)
    pl = O.plan_hankel(x, nu=0, minfolds=3, xy=1, lowring=False)
    assert (pl['N'], pl['in_left'], pl['in_right'], pl['out_left'], pl['out_right']) == (32, 12, 13, 13, 12)
    assert np.allclose(pl['padded_x'], np.logspace(-15, 16, num=32, endpoint=True))
    assert np.allclose(pl['padded_y'], np.logspace(-16, 15, num=32, endpoint=True))


@pytest.mark.parametrize('lowring', [True, False])
def test_execute_direct_pins_fft_semantics(lowring):
    """O(N^2) long-double DFT restatement == FFT-based oracle: sign, conj, dropped Im at DC/Nyquist (lowring=False
    has a complex Nyquist coefficient)."""
    x = np.logspace(-3, 3, num=60, endpoint=False)
    f = 1 / (1 + x**2)**1.5
    pl = O.plan_hankel(x, nu=0, q=1, lowring=lowring)
    if not lowring:
        assert abs(pl['padded_u'][0, -1].imag) > 1e-3
    for extrap in [0, 'log']:
        y, g = O.execute(pl, f, extrap=extrap)
        gd = O.execute_direct(pl, f, extrap=extrap)
        assert np.max(np.abs(gd - g)) < 1e-12 * np.max(np.abs(g))


def test_multi_ell_equals_single(fftlog_golden):
    """tests/test_fftlog.py:107."""
    k, pk = fftlog_golden.inp('k1000'), fftlog_golden.inp('pk1000')
    multi = O.execute(O.plan_power_to_correlation(k, ell=[0, 1, 2, 3, 4]), pk)[1]
    for ell in range(5):
        assert np.allclose(O.execute(O.plan_power_to_correlation(k, ell=ell), pk)[1], multi[ell])
    s = O.execute(O.plan_power_to_correlation(k, ell=0, lowring=False), pk)[0]
    assert np.allclose(s[::-1] * k, 1.)
