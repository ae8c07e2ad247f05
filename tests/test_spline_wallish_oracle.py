"""Spline and Wallish2018 oracles pinned against vectors produced by the reference; CPU emulation of the fused Wallish
kernel against the oracle.  CPU only."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import load_golden
from oracle import spline_oracle as SO
from oracle import wallish_oracle as WO

SPLINE_CASES = list(range(len(load_golden('spline_golden.npz').cases)))


@pytest.mark.parametrize('idx', SPLINE_CASES)
def test_spline_oracle_matches_reference(idx):
    g = load_golden('spline_golden.npz')
    case = g.cases[idx]
    ref = g.data['s{}'.format(idx)]
    with np.errstate(all='ignore'):
        if case['kind'] == 'interp1d':
            f = SO.interpolator1d(g.data[case['x']], g.data[case['y']], interp_x=case['interp_x'], interp_fun=case['interp_fun'],
                                  extrap=case['extrap'], assume_sorted=case.get('assume_sorted', False))
            out = f(g.data[case['xq']], dx=case['dx'])
        else:
            pytest.skip('interpolator-level case: checked on the GPU path')
    assert out.shape == ref.shape and out.dtype == ref.dtype
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    m = ~np.isnan(ref)
    np.testing.assert_allclose(out[m], ref[m], rtol=1e-13, atol=0)


@pytest.mark.parametrize('bc', ['natural', 'clamped'])
def test_thomas_restatement_matches_scipy(bc):
    """The library-free slope solve / PPoly evaluation == scipy.interpolate.CubicSpline (what the reference calls)."""
    from scipy.interpolate import CubicSpline
    g = load_golden('spline_golden.npz')
    for xk, yk, qk in [('x', 'y', 'xq'), ('s_ill', 'var_ill', 'r')]:
        x, y, xq = g.data[xk], g.data[yk], g.data[qk]
        xq = xq[(xq >= x[0]) & (xq <= x[-1])]
        cs = CubicSpline(x, y, axis=0, bc_type=bc)
        s = SO.cubic_spline_slopes(x, y, bc)
        assert np.max(np.abs(s[:-1] - cs.c[2])) <= 1e-13 * np.max(np.abs(cs.c[2]))
        for nu in [0, 1, 2]:
            ref = cs(xq, nu=nu)
            assert np.max(np.abs(SO.cubic_spline_eval(x, y, s, xq, nu=nu) - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize('idx', [0, 1])
def test_wallish_oracle_matches_reference(idx):
    g = load_golden('wallish_golden.npz')
    d = g.data
    pknow, dbg = WO.wallish2018(d['w%d_klin' % idx], d['w%d_pklin' % idx], d['w%d_kout' % idx], d['w%d_pkout' % idx], return_debug=True)
    np.testing.assert_allclose(pknow, d['w%d_pknow' % idx], rtol=1e-12, atol=0)
    np.testing.assert_allclose(dbg['dd_even'][:64], d['w%d_dd_even' % idx], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dbg['even_now'][:128], d['w%d_now_head' % idx][0], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dbg['odd_now'][:128], d['w%d_now_head' % idx][1], rtol=1e-10, atol=1e-14)


def test_wallish_kernel_emulation():
    """tests/emul/emul_wallish.cpp runs the fused kernel's phase functions thread by thread on the CPU."""
    here = os.path.dirname(os.path.abspath(__file__))
    g = load_golden('wallish_golden.npz')
    d = g.data
    N = 4096
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, 'emul_wallish')
        subprocess.run(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(here, 'emul', 'emul_wallish.cpp')], check=True)
        for idx, c0, cap in [(0, 0, 0), (0, 4, 37), (1, 0, 0)]:        # cap: slot array shrunk, several evaluation rounds
            klin, pklin = d['w%d_klin' % idx], d['w%d_pklin' % idx]
            _, dbg = WO.wallish2018(klin, pklin, d['w%d_kout' % idx], d['w%d_pkout' % idx], return_debug=True)
            fin, fout = os.path.join(tmp, 'in.bin'), os.path.join(tmp, 'out.bin')
            kout, pkout = d['w%d_kout' % idx], d['w%d_pkout' % idx]
            np.concatenate([klin, pklin[:, c0], pklin[:, c0 + 1], [float(kout.size)], kout, pkout[:, c0], pkout[:, c0 + 1]]).tofile(fin)
            subprocess.run([exe, fin, fout, str(cap)], check=True)
            r = np.fromfile(fout)
            X, dd = r[:2 * N].reshape(N, 2), r[2 * N:4 * N].reshape(2, N // 2, 2)
            boxes = r[4 * N:4 * N + 8].astype(int).reshape(2, 2, 2)        # [parity, column, (b0, b1)]
            Xnow, pl = r[4 * N + 8:6 * N + 8].reshape(N, 2), r[6 * N + 8:8 * N + 8].reshape(N, 2)
            pknow = r[8 * N + 8:8 * N + 8 + 2 * kout.size].reshape(kout.size, 2)
            nuni, nrounds, nc, lz, rz = r[-5:].astype(int)
            assert np.max(np.abs(X[:256] - d['w%d_dst_head' % idx][:, c0:c0 + 2])) < 1e-14 * np.max(np.abs(X))
            for h, name in enumerate(['dd_even', 'dd_odd']):
                ref = dbg[name][:, c0:c0 + 2]
                assert np.max(np.abs(dd[h] - ref)) < 1e-13 * np.max(np.abs(ref))
            for col in range(2):
                assert boxes[0, col].tolist() == dbg['boxes'][c0 + col, :2].tolist()
                assert boxes[1, col].tolist() == dbg['boxes'][c0 + col, 2:].tolist()
            ref_now = np.empty((N, 2))
            ref_now[::2], ref_now[1::2] = dbg['even_now'][:, c0:c0 + 2], dbg['odd_now'][:, c0:c0 + 2]
            assert np.max(np.abs(Xnow - ref_now)) < 1e-14 * np.max(np.abs(ref_now))
            m = (klin > 1e-2) & (klin < 1.5)
            ref_pl = (np.exp(dbg['kpknow']) / klin[:, None])[:, c0:c0 + 2]
            assert np.max(np.abs(pl[m] / ref_pl[m] - 1)) < 1e-11
            # final stage (spliced clamped spline in the buffer, truncated edges, slot rounds) against the reference's pknow
            ref_pknow = d['w%d_pknow' % idx][:, c0:c0 + 2]
            assert np.array_equal(np.isnan(pknow), np.isnan(ref_pknow))
            ok = ~np.isnan(ref_pknow)
            assert np.max(np.abs(pknow[ok] / ref_pknow[ok] - 1)) < 1e-11
            assert (nrounds == 1 if cap == 0 else nrounds > 10) and nc == lz + m.sum() + rz
            assert nuni > 150          # threads on the uniform stretch of the spliced knots take constant factors
