"""The least-squares / peak-average BAO filters (ref bao_filter.py:289-342, 512-580, 835-909; utils.py:144-272;
eisenstein_hu_nowiggle.py) against vectors produced by the unmodified reference (tools/make_golden.py::make_filters).  The inputs are the
reference's own evaluations (bit-identical arrays), the cosmology numbers come from EHCosmology."""
import numpy as np
import pytest

from conftest import load_golden, reference
from cosmoprimo_b200.utils import LeastSquareSolver
from cosmoprimo_b200.eisenstein_hu import EHCosmology
from cosmoprimo_b200 import bao_filter as B

NAMES = ['h', 'omega_b', 'omega_cdm', 'n_s', 'A_s']


def golden():
    return np.load(__import__('os').path.join(__import__('conftest').GOLDEN_DIR, 'filters_golden.npz'))


def cosmologies(d):
    return EHCosmology(**dict(zip(NAMES, d['par']))), EHCosmology(**dict(zip(NAMES, d['par_fid'])))


class TablePk(object):
    """Interpolator duck type returning the stored evaluations of the reference's interpolator (bit-identical inputs)."""

    def __init__(self, d, device=False):
        self.extrap_kmin, self.extrap_kmax = float(d['extrap_kmin']), float(d['extrap_kmax'])
        self._k, self._pk, self._device = d['k'], d['pk'], device

    def __call__(self, k):
        assert np.allclose(k, self._k, rtol=1e-14)
        if self._device:
            import torch
            return torch.from_numpy(self._pk).cuda()
        return self._pk


class TableXi(object):

    def __init__(self, d, device=False):
        self.extrap_smin, self.extrap_smax = float(d['extrap_smin']), float(d['extrap_smax'])
        self._s, self._xi, self._device = d['s'], d['xi'], device

    def __call__(self, s):
        assert np.allclose(s, self._s, rtol=1e-14)
        if self._device:
            import torch
            return torch.from_numpy(self._xi).cuda()
        return self._xi


def test_eh_cosmology_matches_stored_reference_numbers():
    d = golden()
    cosmo, fid = cosmologies(d)
    assert abs(cosmo.rs_drag / float(d['rs_drag']) - 1.) < 1e-14 and abs(fid.rs_drag / float(d['rs_drag_fid']) - 1.) < 1e-14
    np.testing.assert_allclose(cosmo.pk_nowiggle(d['k']), d['pk_nowiggle'], rtol=1e-13)


def test_least_square_solver_against_reference():
    """utils.py:144-272: same parameters, model and chi2 as the reference's solver, with and without constraints / inverse."""
    ref = reference('utils').LeastSquareSolver
    k = np.geomspace(1e-3, 1., 300)
    rng = np.random.default_rng(0)
    data = 1. + 0.05 * np.sin(60. * k) * np.exp(-5. * k) + 0.01 * rng.standard_normal((7, 1))
    gradient = np.array([k**(i - 2) for i in range(6)])
    cg = np.column_stack([gradient[..., 0], gradient[..., 1] - gradient[..., 0], gradient[..., -1], gradient[..., -2] - gradient[..., -1]])
    con = np.column_stack([data[..., 0], data[..., 1] - data[..., 0], data[..., -1], data[..., -2] - data[..., -1]])
    for kw, ckw in [(dict(precision=k**2, constraint_gradient=cg), dict(constraint=con)), (dict(precision=k**2), {}), (dict(), {}),
                    (dict(precision=np.diag(k**2)), {})]:
        for inverse in [True, False]:
            r, o = ref(gradient, compute_inverse=inverse, **kw), LeastSquareSolver(gradient, compute_inverse=inverse, **kw)
            pr, po = r(data, **ckw), o(data, **ckw)
            assert po.shape == pr.shape
            np.testing.assert_allclose(o.model(), r.model(), rtol=1e-11)
            np.testing.assert_allclose(o.chi2(), r.chi2(), rtol=1e-6, atol=1e-20)
    one = LeastSquareSolver(np.ones(4))                      # the docstring example of the reference (utils.py:154-160)
    assert one(2 * np.ones(4)) == 2.0 and np.array_equal(one.model(), 2. * np.ones(4)) and one.chi2() == 0.


@pytest.mark.gpu
@pytest.mark.parametrize('device', [False, True])
def test_ehpoly_filter(device):
    d = golden()
    cosmo, fid = cosmologies(d)
    filt = B.PowerSpectrumBAOFilter(TablePk(d, device), engine='ehpoly', cosmo=cosmo)
    assert type(filt) is B.EHNoWigglePolyPowerSpectrumBAOFilter and np.allclose(filt.k, d['k'], rtol=1e-14)
    out = filt.pknow.cpu().numpy() if device else filt.pknow
    assert out.shape == d['ehpoly_pknow'].shape == (1024, 4)
    # the polynomial basis k^-2 .. k^3 is ill conditioned (normal matrix cond ~1e10): rounding differences of the solve show up at ~1e-9
    np.testing.assert_allclose(out, d['ehpoly_pknow'], rtol=2e-8)
    filt = B.PowerSpectrumBAOFilter(TablePk(d, device), engine='ehpoly_cuda', cosmo=cosmo, cosmo_fid=fid, krange=(2e-3, 0.8), rescale_krange=True)
    out = filt.pknow.cpu().numpy() if device else filt.pknow
    np.testing.assert_allclose(out, d['ehpoly2_pknow'], rtol=2e-8)
    with pytest.raises(ValueError):
        B.PowerSpectrumBAOFilter(TablePk(d), engine='ehpoly')          # no cosmology given


@pytest.mark.gpu
@pytest.mark.parametrize('device', [False, True])
def test_peakaverage_filter(device):
    d = golden()
    cosmo, fid = cosmologies(d)
    filt = B.PowerSpectrumBAOFilter(TablePk(d, device), engine='peakaverage', cosmo=cosmo, cosmo_fid=fid)
    assert type(filt) is B.PeakAveragePowerSpectrumBAOFilter
    # the fiducial peak positions (scipy.signal.find_peaks on the fiducial wiggles) are integers: exact
    assert np.array_equal(np.array(filt.pad_peaks), d['peakaverage_pad'])
    np.testing.assert_allclose(filt.k_peaks[0], d['peakaverage_k_peaks0'], rtol=1e-14)
    np.testing.assert_allclose(filt.k_peaks[1], d['peakaverage_k_peaks1'], rtol=1e-14)
    out = filt.pknow.cpu().numpy() if device else filt.pknow
    assert out.shape == (1024, 4)
    np.testing.assert_allclose(out, d['peakaverage_pknow'], rtol=1e-10)
    # wiggles oscillate around one
    w = (d['pk'] / out)[(d['k'] > 0.02) & (d['k'] < 0.3)]
    assert 0.9 < w.min() < 1. < w.max() < 1.1
    with pytest.raises(ValueError):
        B.PowerSpectrumBAOFilter(TablePk(d), engine='peakaverage', cosmo=cosmo)      # cosmo_fid is mandatory (ref:521-524)


@pytest.mark.gpu
@pytest.mark.parametrize('device', [False, True])
def test_kirkby2013_filter(device):
    d = golden()
    cosmo, fid = cosmologies(d)
    filt = B.CorrelationFunctionBAOFilter(TableXi(d, device), engine='kirkby2013', cosmo=cosmo)
    assert type(filt) is B.Kirkby2013CorrelationFunctionBAOFilter and np.allclose(filt.s, d['s'], rtol=1e-14)
    out = filt.xinow.cpu().numpy() if device else filt.xinow
    scale = np.max(np.abs(d['kirkby_xinow']), axis=0)
    assert np.max(np.abs(out - d['kirkby_xinow']) / scale) < 1e-10
    filt = B.CorrelationFunctionBAOFilter(TableXi(d, device), engine='kirkby2013_cuda', cosmo=cosmo, cosmo_fid=fid, srange_left=(45., 80.), srange_right=(155., 195.))
    out = filt.xinow.cpu().numpy() if device else filt.xinow
    assert np.max(np.abs(out - d['kirkby2_xinow']) / scale) < 1e-10
    with pytest.raises(ValueError):
        B.CorrelationFunctionBAOFilter(TableXi(d), engine='nope')


@pytest.mark.gpu
def test_new_filters_register_in_reference():
    """The reference's factories build our classes by name once they are registered (ref bao_filter.py:22-31, 691-700, 912-933), with the
    reference's own Cosmology objects standing in for EHCosmology."""
    Cosmology = reference().Cosmology          # puts baseline/_ref on the path
    refb = B.register_in_reference()
    d = golden()
    cosmo = Cosmology(m_ncdm=None, engine='eisenstein_hu', **dict(zip(NAMES, d['par'])))
    fid = Cosmology(m_ncdm=None, engine='eisenstein_hu', **dict(zip(NAMES, d['par_fid'])))
    filt = refb.PowerSpectrumBAOFilter(TablePk(d), engine='peakaverage_cuda', cosmo=cosmo, cosmo_fid=fid)
    np.testing.assert_allclose(filt.pknow, d['peakaverage_pknow'], rtol=1e-10)
    filt = refb.PowerSpectrumBAOFilter(TablePk(d), engine='ehpoly_cuda', cosmo=cosmo)
    np.testing.assert_allclose(filt.pknow, d['ehpoly_pknow'], rtol=2e-8)
    filt = refb.CorrelationFunctionBAOFilter(TableXi(d), engine='kirkby2013_cuda', cosmo=cosmo)
    assert np.max(np.abs(filt.xinow - d['kirkby_xinow']) / np.max(np.abs(d['kirkby_xinow']), axis=0)) < 1e-10
