"""Host side of the cuda engine (plan tables, grids, pad, registry, shape rules) against the reference's golden
vectors.  CPU only: nothing here launches a kernel."""
import numpy as np
import pytest

from conftest import load_golden, Golden
from cosmoprimo_b200 import fftlog as F

CASES = list(range(len(load_golden().cases)))


def build(golden, idx):
    case = golden.cases[idx]
    obj = getattr(F, case['cls'])(golden.inp(case['grid']), **Golden.ctor_kwargs(case))
    if case['inv']:
        obj.inv()
    return obj


@pytest.mark.parametrize('idx', CASES)
def test_plan_matches_reference(fftlog_golden, idx):
    case = fftlog_golden.cases[idx]
    obj = build(fftlog_golden, idx)
    assert (obj.padded_size, obj.size, obj.nparallel) == (case['N'], case['n'], case['P'])
    keep = case['callkw'].get('keep_padding', False)
    y = obj.padded_y if keep else obj.y
    if not obj.inparallel: y = y[0]
    np.testing.assert_allclose(y, fftlog_golden.get(idx, 'y'), rtol=1e-14, atol=0)
    if case['tables']:
        np.testing.assert_allclose(obj.padded_prefactor, fftlog_golden.get(idx, 'pre'), rtol=1e-13, atol=0)
        np.testing.assert_allclose(obj.padded_u, fftlog_golden.get(idx, 'u'), rtol=1e-13, atol=0)
        np.testing.assert_allclose(obj.padded_postfactor, fftlog_golden.get(idx, 'post'), rtol=1e-13, atol=0)
        np.testing.assert_allclose(obj.padded_x, fftlog_golden.get(idx, 'padded_x'), rtol=1e-14, atol=0)
        np.testing.assert_allclose(obj.padded_y, fftlog_golden.get(idx, 'padded_y'), rtol=1e-14, atol=0)
        assert obj.padded_postfactor.dtype == fftlog_golden.get(idx, 'post').dtype


def test_pad():
    """Mirror of the reference's test_pad (tests/test_fftlog.py:26-53)."""
    a = b = np.ones((6, 6))
    padded_a = np.zeros((13, 6))
    padded_a[3: 9, :] = 1
    padded_b = np.ones((6, 13))
    c = np.array([(i + 1) * np.logspace(-3, 3, num=6, endpoint=False) for i in range(3)]).T
    padded_c = np.array([(i + 1) * np.logspace(-12, 12, num=24, endpoint=False) for i in range(3)]).T
    assert np.allclose(F.pad(a, (3, 4), extrap=0, axis=0), padded_a)
    assert np.allclose(F.pad(b, (4, 3), extrap='edge', axis=1), padded_b)
    assert np.allclose(F.pad(c, (9, 9), extrap='log', axis=0), padded_c)
    assert np.allclose(F.pad([1., 2., 4., 8.], (2, 3), extrap=(0, 'log')), [0, 0, 1, 2, 4, 8, 16, 32, 64])

    x = np.logspace(-3, 3, num=7, endpoint=True)
    padded_x = np.logspace(-15, 16, num=32, endpoint=True)
    y = np.logspace(-3, 3, num=7, endpoint=True)
    padded_y = np.logspace(-16, 15, num=32, endpoint=True)
    fftlog = F.HankelTransform(x, minfolds=3, xy=1, lowring=False)
    assert np.allclose(fftlog.padded_x, padded_x)
    assert np.allclose(fftlog.padded_y, padded_y)
    assert np.allclose(F.pad(x, (fftlog.padded_size_in_left, fftlog.padded_size_in_right), extrap='log'), padded_x)
    assert np.allclose(F.pad(y, (fftlog.padded_size_out_left, fftlog.padded_size_out_right), extrap='log'), padded_y)
    assert np.allclose(fftlog.padded_x[0, fftlog.padded_size_in_left: fftlog.padded_size_in_left + fftlog.size], x)
    assert np.allclose(fftlog.padded_y[0, fftlog.padded_size_out_left: fftlog.padded_size_out_left + fftlog.size], y)


def test_sizes():
    """N = smallest power of two >= n * minfolds (fftlog.py:149-150), not the docstring's strict inequality."""
    for n, N in [(1000, 2048), (1024, 2048), (2048, 4096), (4096, 8192), (60, 128)]:
        assert F.PowerToCorrelation(np.geomspace(1e-3, 1e1, n)).padded_size == N


def test_engine_registry():
    k = np.geomspace(1e-3, 1e1, 64)
    f = F.PowerToCorrelation(k, engine='cuda')
    assert isinstance(f._engine, F.CudaFFTEngine) and f._engine.size == 128 and f._engine.nparallel == 1
    assert F.get_fft_engine(f._engine) is f._engine                      # instances pass through (fftlog.py:663)
    assert f._engine == F.CudaFFTEngine(128) and hash(f._engine) == hash(F.CudaFFTEngine(128))
    with pytest.raises(ValueError):
        F.get_fft_engine('nope', size=128)                                # unknown engine (fftlog.py:662)
    with pytest.raises(ValueError):
        F.PowerToCorrelation(k, engine='numpy')                           # CPU engines live in the reference
    with pytest.raises(TypeError):
        F.PowerToCorrelation(k, engine=object())                          # no silent unfused/CPU path
    import os
    before = os.environ.get('OMP_NUM_THREADS')
    F.CudaFFTEngine(128, nthreads=7)
    assert os.environ.get('OMP_NUM_THREADS') == before                    # unlike BaseFFTEngine (fftlog.py:529-531)


def test_kernel_equality():
    assert F.SphericalBesselJKernel(2) == F.SphericalBesselJKernel(2)
    assert F.SphericalBesselJKernel(2) != F.SphericalBesselJKernel(0)
    assert F.SphericalBesselJKernel(2) != F.BesselJKernel(2)
    assert F.TophatSqKernel(3) == F.TophatSqKernel(ndim=3) and F.TophatSqKernel(3) != F.TophatSqKernel(1)
    assert len({F.GaussianKernel(), F.GaussianKernel(), F.GaussianSqKernel()}) == 2


def test_check_level():
    k = np.geomspace(1e-3, 1e1, 64)
    F.PowerToCorrelation(k, check_level=1)
    with pytest.raises(ValueError):
        F.PowerToCorrelation(np.linspace(1e-3, 1e1, 64), check_level=1)
    with pytest.raises(ValueError):
        F.FFTlog(k, [F.BesselJKernel(0), F.BesselJKernel(1)], q=[1.], check_level=1)


def test_plan_cache_returns_private_copies():
    k = np.geomspace(1e-3, 1e1, 64)
    a = F.PowerToCorrelation(k)
    b = F.PowerToCorrelation(k)
    assert np.array_equal(a.padded_prefactor, b.padded_prefactor)
    a.padded_prefactor *= 2.
    assert not np.array_equal(a.padded_prefactor, b.padded_prefactor)
    c = F.PowerToCorrelation(k)
    assert np.array_equal(c.padded_prefactor, b.padded_prefactor)


def test_bad_input_shape():
    k = np.geomspace(1e-3, 1e1, 64)
    with pytest.raises(ValueError):
        F.PowerToCorrelation(k)(np.ones(63))
    with pytest.raises(ValueError):
        F.PowerToCorrelation(k, ell=[0, 2])(np.ones((5, 64)))             # (5,) does not broadcast against (2,)
