"""Eisenstein & Hu generator: golden vectors from the reference's engine (tools/make_golden.py::make_eh) against the numpy
generator (cosmoprimo_b200/synthetic.py), the point functions of the CUDA kernel run on the CPU (tests/emul/emul_eh.cpp) and,
on a GPU, the kernel itself through the C ABI."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from cosmoprimo_b200 import synthetic as S

RTOL = 1e-12   # fp64; libm / CUDA transcendental functions differ by a few ulp, ~20 of them enter each P(k)


def golden():
    d = np.load(os.path.join(GOLDEN_DIR, 'eh_golden.npz'))
    par = {n: d['par_' + n] for n in S.PARAM_NAMES}
    return d, par


def flat_inputs(d, par):
    """every (cosmology, redshift) pair as one row: params (B*nz, 5), z (B*nz,)"""
    nz = d['z'].size
    cols = [par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], 1e-10 * np.exp(par['logA'])]
    params = np.repeat(np.stack(cols, axis=-1), nz, axis=0)
    z = np.tile(d['z'], par['h'].size)
    return params, z


def test_numpy_generator_matches_reference():
    d, par = golden()
    assert abs(S.omega_radiation() / float(d['omega_r']) - 1) < 1e-14
    for iz, z in enumerate(d['z']):
        np.testing.assert_allclose(S.eh_pk(d['k'], par, z=z), d['pk'][:, iz], rtol=1e-13)
    Om0 = (par['omega_b'] + par['omega_cdm']) / par['h']**2
    np.testing.assert_allclose(S.growth_rate(0.5, Om0, par['h']), d['derived'][:, 1, 3], rtol=1e-13)


def test_point_functions_on_cpu():
    d, par = golden()
    params, z = flat_inputs(d, par)
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_eh')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_eh.cpp')], check=True)
        head = np.array([params.shape[0], d['k'].size, float(d['T_cmb']), float(d['omega_r']), float(d['k_pivot'])])
        np.concatenate([head, params.ravel(), z, d['k']]).tofile(os.path.join(tmp, 'in.bin'))
        subprocess.run([exe, os.path.join(tmp, 'in.bin'), os.path.join(tmp, 'out.bin')], check=True)
        out = np.fromfile(os.path.join(tmp, 'out.bin'))
    B, nk = params.shape[0], d['k'].size
    np.testing.assert_allclose(out[:B * nk].reshape(-1, d['z'].size, nk), d['pk'], rtol=RTOL)
    np.testing.assert_allclose(out[B * nk:].reshape(-1, d['z'].size, 4), d['derived'], rtol=RTOL)


@pytest.mark.gpu
def test_cuda_generator_matches_reference():
    torch = pytest.importorskip('torch')
    from cosmoprimo_b200.eisenstein_hu import EisensteinHu
    d, par = golden()
    params, z = flat_inputs(d, par)
    eh = EisensteinHu(*params.T[:4], A_s=params[:, 4])
    assert abs(eh.omega_r / float(d['omega_r']) - 1) < 1e-14
    pk = eh.pk(d['k'], z=z)
    assert isinstance(pk, torch.Tensor) and pk.is_cuda and tuple(pk.shape) == (params.shape[0], d['k'].size)
    np.testing.assert_allclose(pk.cpu().numpy().reshape(d['pk'].shape), d['pk'], rtol=RTOL)
    host = eh.pk(d['k'], z=z, on_device=False)
    assert isinstance(host, np.ndarray) and np.array_equal(host, pk.cpu().numpy())
    np.testing.assert_allclose(eh.derived(z=z).reshape(d['derived'].shape), d['derived'], rtol=RTOL)
    # Kaiser multipoles = numpy generator's, z = None means z = 0, logA constructor
    eh0 = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    Om0 = (par['omega_b'] + par['omega_cdm']) / par['h']**2
    ref = S.kaiser_multipoles(S.eh_pk(d['k'], par, z=0.), S.growth_rate(0., Om0, par['h']))
    np.testing.assert_allclose(eh0.pk(d['k'], kaiser=True).cpu().numpy(), ref, rtol=RTOL)
    # seeded batch against the numpy generator, one redshift per row, nk = 2048
    B = 300
    par2 = S.lhs_cosmologies(B, seed=7)
    zz = np.random.default_rng(7).uniform(0., 3., B)
    k = np.geomspace(1e-5, 1e2, 2048)
    eh2 = EisensteinHu(par2['h'], par2['omega_b'], par2['omega_cdm'], par2['n_s'], logA=par2['logA'])
    np.testing.assert_allclose(eh2.pk(k, z=zz).cpu().numpy(), S.eh_pk(k, par2, z=zz), rtol=RTOL)
    # redshift grid per cosmology: (B, nz, nk), one transfer-function evaluation per cosmology
    eh9 = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    grid = eh9.pk(d['k'], z=d['z'][None, :])
    assert tuple(grid.shape) == (par['h'].size, d['z'].size, d['k'].size)
    np.testing.assert_allclose(grid.cpu().numpy(), d['pk'], rtol=RTOL)
    np.testing.assert_allclose(eh9.derived(z=d['z'][None, :]), d['derived'], rtol=RTOL)
    assert tuple(eh9.pk(d['k'], z=d['z'][None, :], kaiser=True).shape) == (par['h'].size, d['z'].size, 3, d['k'].size)


@pytest.mark.gpu
def test_generator_feeds_fftlog_on_device():
    """generator -> FFTLog with no host copy of the spectra: same multipoles as host-generated input"""
    torch = pytest.importorskip('torch')
    from cosmoprimo_b200.eisenstein_hu import EisensteinHu
    from cosmoprimo_b200.fftlog import PowerToCorrelation
    B, n = 40, 2048
    par = S.lhs_cosmologies(B, seed=3)
    k = np.geomspace(1e-5, 1e2, n)
    eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    fun = eh.pk(k, z=0.5, kaiser=True)
    fftlog = PowerToCorrelation(k, ell=[0, 2, 4])
    s, xi = fftlog(fun)
    s2, xi2 = fftlog(fun.cpu().numpy())
    assert isinstance(xi, torch.Tensor) and np.array_equal(xi.cpu().numpy(), xi2)
