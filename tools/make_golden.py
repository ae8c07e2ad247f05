"""
Generate the golden vectors of tests/golden/ by running the UNMODIFIED reference (cosmodesi/cosmoprimo, mounted
read-only at /root/reference) with its numpy engine.  The reference cannot travel to the GPU box, the vectors can.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

(The reference imports through the 3-line dist-info shim in tools/refshim)
Outputs: tests/golden/fftlog_golden.npz (+ spline / wallish files written by the other functions below).
"""

import os
import sys
import json

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'refshim'))

import numpy as np

import cosmoprimo
from cosmoprimo import fftlog as ref
from cosmoprimo.fiducial import DESI

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def desi_pk(k, z=0.):
    cosmo = DESI(engine='eisenstein_hu')
    return cosmo.get_fourier().pk_interpolator()(k, z=z)


def class_table_pk(k, fn):
    """One of the reference's CLASS P(k) tables (real spectrum with BAO), log-log interpolated onto ``k``."""
    kt, pt = np.loadtxt(fn, unpack=True)[:2]
    return np.exp(np.interp(np.log(k), np.log(kt), np.log(pt)))


def make_fftlog():
    arrays, cases = {}, []

    def add_input(name, arr):
        arrays['in_' + name] = np.asarray(arr)

    def add_case(cls, grid, fun, ckw=None, callkw=None, tables=False, inv=False, tag=''):
        ckw, callkw = dict(ckw or {}), dict(callkw or {})
        run_kw = {k: (np.asarray(v) if k == 'q' and isinstance(v, list) else v) for k, v in ckw.items()}   # q lists must be arrays (fftlog.py:318)
        obj = getattr(ref, cls)(arrays['in_' + grid], engine='numpy', **run_kw)
        if inv:
            # transform forward first (its output is the inverse's input), then invert in place
            y0, g0 = obj(arrays['in_' + fun], **callkw)
            obj.inv()
            f_in = g0
            idx = len(cases)
            arrays['c{}_fun'.format(idx)] = f_in
        else:
            f_in = arrays['in_' + fun]
        y, g = obj(f_in, **callkw)
        idx = len(cases)
        arrays['c{}_y'.format(idx)] = y
        arrays['c{}_g'.format(idx)] = g
        if tables:
            arrays['c{}_pre'.format(idx)] = obj.padded_prefactor
            arrays['c{}_u'.format(idx)] = obj.padded_u
            arrays['c{}_post'.format(idx)] = obj.padded_postfactor
            arrays['c{}_padded_x'.format(idx)] = obj.padded_x
            arrays['c{}_padded_y'.format(idx)] = obj.padded_y
        jkw = {k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in callkw.items()}
        cases.append(dict(cls=cls, grid=grid, fun=fun, ckw=ckw, callkw=jkw, tables=tables, inv=inv, tag=tag,
                          N=int(obj.padded_size), n=int(obj.size), P=int(obj.nparallel)))

    # ---- inputs ------------------------------------------------------------------------------------------------
    for n in [1000, 1024, 2048, 4096]:
        k = np.geomspace(1e-5, 1e2, n) if n != 1000 else np.logspace(-5, 2, 1000)   # 1000: grid of the reference tests
        add_input('k{}'.format(n), k)
        add_input('pk{}'.format(n), desi_pk(k))
    k = arrays['in_k2048']
    pk = arrays['in_pk2048']
    # Kaiser multipoles (BASELINE config 2), f = Omega_m(z=0.5)^0.55
    f = 0.76
    add_input('pkmulti2048', np.array([(1 + 2 * f / 3 + f**2 / 5) * pk, (4 * f / 3 + 4 * f**2 / 7) * pk, (8 * f**2 / 35) * pk]))
    rng = np.random.default_rng(42)
    scales = 1. + 0.5 * rng.uniform(size=(4, 1))
    add_input('pkbatch1000', arrays['in_pk1000'][None, :] * np.linspace(1., 3., 5)[:, None])
    add_input('pkbatch2048_b1', (arrays['in_pk2048'][None, :] * scales)[:, None, :])                    # (4,1,n)
    add_input('pkbatch2048_b3', arrays['in_pkmulti2048'][None, :, :] * scales[:, :, None])              # (4,3,n)
    fid = '/root/reference/cosmoprimo/tests/fiducial'
    add_input('pkclass2048', np.array([class_table_pk(k[(k > 2e-5) & (k < 50)], os.path.join(fid, 'abacus_cosm000_CLASSv3.1.1.00_z{}_pk.dat'.format(i))) for i in (1, 3)]))
    add_input('kclass2048', k[(k > 2e-5) & (k < 50)])
    x60 = np.logspace(-3, 3, num=60, endpoint=False)
    add_input('x60', x60)
    add_input('f60', 1 / (1 + x60**2)**1.5)
    add_input('x7', np.logspace(-3, 3, num=7, endpoint=True))
    add_input('f7', 1 / (1 + arrays['in_x7']**2)**1.5)

    # ---- cases -------------------------------------------------------------------------------------------------
    # A: analytic Hankel pair of the reference tests (test_fftlog.py:56-89)
    add_case('HankelTransform', 'x60', 'f60', dict(nu=0, q=1, lowring=True), dict(extrap='log'), tables=True, tag='hankel analytic')
    add_case('HankelTransform', 'x60', 'f60', dict(nu=0, q=1, lowring=True), dict(extrap='log'), inv=True, tag='hankel inv')
    add_case('HankelTransform', 'x7', 'f7', dict(nu=0, q=1, minfolds=3, xy=1, lowring=False), dict(extrap='log'), tables=True, tag='odd padding (test_pad)')
    add_case('HankelTransform', 'x7', 'f7', dict(nu=[0, 1], q=1, minfolds=3, lowring=True), dict(extrap='edge', keep_padding=True), tag='odd padding, P=2, keep')
    # B: n = 1000 (N = 2048), the grid of test_power_to_correlation
    for ell in range(5):
        add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=ell, lowring=True, complex=False), tables=(ell == 2), tag='P2xi ell')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=[0, 1, 2, 3, 4], lowring=True, q=0, complex=False), tag='multi ell')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=0, lowring=False), tag='lowring False')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=0, lowring=False, xy=2.5), tag='xy')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=2, q=0.5), tag='q')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=[0, 2], q=[0., 0.5]), tag='q list')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=[0, 1, 2], complex=True), tables=True, tag='complex post')
    for extrap in ['log', 'edge', (0, 'log'), ('edge', 1e3), 2.5]:
        add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=0), dict(extrap=extrap), tag='extrap')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=0), dict(keep_padding=True), tag='keep_padding')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=[0, 2]), dict(extrap='log', keep_padding=True), tag='keep_padding log P=2')
    add_case('PowerToCorrelation', 'k1000', 'pkbatch1000', dict(ell=0), tag='batch (5,n)')
    add_case('PowerToCorrelation', 'k1000', 'pk1000', dict(ell=1), inv=True, tag='inv')
    add_case('TophatVariance', 'k1000', 'pk1000', dict(lowring=True), tables=True, tag='sigma_r')
    add_case('GaussianVariance', 'k1000', 'pk1000', dict(), tag='gaussian')
    add_case('HankelTransform', 'k1000', 'pk1000', dict(nu=[0, 2], q=1), dict(extrap='log'), tag='hankel P=2')
    # C: n = 1024 (N = 2048) and n = 2048 (N = 4096), the BASELINE grids
    add_case('PowerToCorrelation', 'k1024', 'pk1024', dict(ell=0), tables=True, tag='config 1')
    add_case('TophatVariance', 'k1024', 'pk1024', dict(), tag='sigma_r 1024')
    add_case('PowerToCorrelation', 'k2048', 'pk2048', dict(ell=0), tables=True, tag='nk=2048 ell=0')
    add_case('PowerToCorrelation', 'k2048', 'pkmulti2048', dict(ell=[0, 2, 4]), tables=True, tag='config 2 (3,n)')
    add_case('PowerToCorrelation', 'k2048', 'pk2048', dict(ell=[0, 2, 4]), tag='config 2 broadcast (n,)')
    add_case('PowerToCorrelation', 'k2048', 'pkbatch2048_b1', dict(ell=[0, 2, 4]), tag='config 2 broadcast (4,1,n)')
    add_case('PowerToCorrelation', 'k2048', 'pkbatch2048_b3', dict(ell=[0, 2, 4]), tag='config 2 (4,3,n)')
    add_case('PowerToCorrelation', 'k2048', 'pk2048', dict(ell=0), dict(extrap='log', keep_padding=True), tag='nk=2048 full variant')
    add_case('TophatVariance', 'k2048', 'pk2048', dict(), tag='config 3')
    add_case('PowerToCorrelation', 'k2048', 'pk2048', dict(ell=0), inv=True, tag='config 5 round trip')
    add_case('PowerToCorrelation', 'kclass2048', 'pkclass2048', dict(ell=0), tag='CLASS tables, n=1919')
    # D: n = 4096 (N = 8192)
    add_case('PowerToCorrelation', 'k4096', 'pk4096', dict(ell=0), tag='nk=4096')
    add_case('CorrelationToPower', 'k4096', 'pk4096', dict(ell=2, complex=True), dict(extrap='edge'), tag='xi2P complex 4096')

    arrays['manifest'] = np.array(json.dumps(cases))
    os.makedirs(GOLDEN, exist_ok=True)
    fn = os.path.join(GOLDEN, 'fftlog_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({} cases, {:.2f} MB)'.format(fn, len(cases), os.path.getsize(fn) / 1e6))


def make_spline():
    """Interpolator1D (jax.py:134-196) and the interpolator-level callers (interpolator.py) run by the reference."""
    from cosmoprimo.jax import Interpolator1D
    from cosmoprimo.interpolator import PowerSpectrumInterpolator1D
    arrays, cases = {}, []
    rng = np.random.default_rng(7)
    x = np.sort(rng.uniform(0.1, 10., 200))
    y = np.sin(x)[:, None] * rng.uniform(1., 2., (200, 7)) + 3.
    xq = np.concatenate([[0.05, x[0], x[-1], 11.], rng.uniform(0.1, 10., 60)])
    arrays.update(x=x, y=y, xq=xq)
    for ix in ['lin', 'log']:
        for ifun in ['lin', 'log']:
            for ex in [False, True]:
                interp = Interpolator1D(x, y, interp_x=ix, interp_fun=ifun, extrap=ex)
                with np.errstate(all='ignore'):
                    for dx in ([0, 1, 2] if (ix, ifun) == ('lin', 'lin') else [0]):
                        arrays['s{}'.format(len(cases))] = interp(xq, dx=dx)
                        cases.append(dict(kind='interp1d', x='x', y='y', xq='xq', interp_x=ix, interp_fun=ifun, extrap=ex, dx=dx))
    # unsorted knots, N-D values, float32 queries, an all-NaN column
    perm = rng.permutation(x.size)
    y3 = y[:, :6].reshape(200, 2, 3).copy()
    y3[:, 1, 2] = np.nan
    arrays.update(x_perm=x[perm], y_perm=y3[perm], xq32=xq.astype('f4').reshape(8, 8))
    interp = Interpolator1D(arrays['x_perm'], arrays['y_perm'])
    arrays['s{}'.format(len(cases))] = interp(arrays['xq32'])
    cases.append(dict(kind='interp1d', x='x_perm', y='y_perm', xq='xq32', interp_x='lin', interp_fun='lin', extrap=False, dx=0))
    # the sigma(r) spline: linear abscissa on a log grid spanning 9 decades (interpolator.py:289)
    s_ = np.geomspace(1e-2, 1e7, 2048)
    var = (1. / (1. + s_)**1.5)[:, None] * np.array([1., 2., 0.5])
    r = np.linspace(1., 20., 10)
    arrays.update(s_ill=s_, var_ill=var, r=r)
    arrays['s{}'.format(len(cases))] = Interpolator1D(s_, var, assume_sorted=True)(r)
    cases.append(dict(kind='interp1d', x='s_ill', y='var_ill', xq='r', interp_x='lin', interp_fun='lin', extrap=False, dx=0, assume_sorted=True))
    # interpolator-level callers
    ktab = np.geomspace(1e-4, 50., 300)
    pk1 = desi_pk(ktab)
    pk3 = pk1[:, None] * np.array([1., 0.5, 2.])
    keval = np.concatenate([[5e-8, 1e-7, 1e2, 2e2], np.geomspace(1e-7, 1e2, 200)])
    arrays.update(ktab=ktab, pk1=pk1, pk3=pk3, keval=keval)
    for name in ['pk1', 'pk3']:
        interp = PowerSpectrumInterpolator1D(ktab, arrays[name])
        idx = len(cases)
        arrays['s{}'.format(idx)] = interp(keval)
        arrays['s{}_sigma_r'.format(idx)] = interp.sigma_r(r)
        arrays['s{}_sigma8'.format(idx)] = np.asarray(interp.sigma8())
        xi = interp.to_xi()
        seval = np.geomspace(xi.smin * 1.01, xi.smax * 0.99, 50)
        arrays['s{}_seval'.format(idx)] = seval
        arrays['s{}_xi'.format(idx)] = xi(seval)
        arrays['s{}_xi_s'.format(idx)] = xi.s
        if name == 'pk1':   # the reference's 1-D to_pk does not transpose multi-column xi (interpolator.py:1212): single column only
            back = xi.to_pk()
            arrays['s{}_pk_back'.format(idx)] = back(np.geomspace(1e-2, 10., 30))
        cases.append(dict(kind='pk1d', k='ktab', pk=name, keval='keval'))
    arrays['manifest'] = np.array(json.dumps(cases))
    fn = os.path.join(GOLDEN, 'spline_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({} cases, {:.2f} MB)'.format(fn, len(cases), os.path.getsize(fn) / 1e6))


def make_wallish():
    """Wallish2018PowerSpectrumBAOFilter (bao_filter.py:345-431) run by the reference on EH (LHS) and CLASS spectra."""
    from cosmoprimo.interpolator import PowerSpectrumInterpolator1D
    from cosmoprimo.bao_filter import PowerSpectrumBAOFilter
    from scipy import fftpack
    sys.path.insert(0, ROOT)
    from cosmoprimo_b200 import synthetic
    arrays, cases = {}, []
    ktab = np.geomspace(1e-5, 1e2, 512)
    pk_eh = synthetic.eh_pk(ktab, synthetic.lhs_cosmologies(6, seed=42)).T             # (512, 6)
    fid = '/root/reference/cosmoprimo/tests/fiducial'
    kc, pc = np.loadtxt(os.path.join(fid, 'abacus_cosm000_CLASSv3.1.1.00_z1_pk.dat'), unpack=True)[:2]
    pc3 = np.loadtxt(os.path.join(fid, 'abacus_cosm000_CLASSv3.1.1.00_z3_pk.dat'), unpack=True)[1]
    sel = slice(None, None, 8)
    arrays.update(ktab_eh=ktab, pk_eh=pk_eh, ktab_class=kc[sel], pk_class=np.array([pc[sel], pc3[sel]]).T)
    for name in ['eh', 'class']:
        interp = PowerSpectrumInterpolator1D(arrays['ktab_' + name], arrays['pk_' + name])
        filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
        klin = np.linspace(interp.extrap_kmin, 2., 4096)
        pklin = interp(klin)
        idx = len(cases)
        arrays['w{}_klin'.format(idx)] = klin
        arrays['w{}_pklin'.format(idx)] = pklin
        arrays['w{}_kout'.format(idx)] = filt.k
        arrays['w{}_pkout'.format(idx)] = filt.pk
        arrays['w{}_pknow'.format(idx)] = filt.pknow
        arrays['w{}_dd_even'.format(idx)] = filt._dd_even[:64]      # head of the second derivatives (where the boxes are)
        arrays['w{}_dd_odd'.format(idx)] = filt._dd_odd[:64]
        # the pre-cut DST coefficients must be recomputed: _even/_odd are views overwritten in place
        dst = fftpack.dst(np.log(klin[:, None] * pklin), type=2, axis=0, norm='ortho')
        arrays['w{}_dst_head'.format(idx)] = dst[:256]
        arrays['w{}_now_head'.format(idx)] = np.stack([filt._even_now[:128], filt._odd_now[:128]])
        cases.append(dict(ktab='ktab_' + name, pk='pk_' + name, ncols=int(pklin.shape[1])))
    arrays['manifest'] = np.array(json.dumps(cases))
    fn = os.path.join(GOLDEN, 'wallish_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({} cases, {:.2f} MB)'.format(fn, len(cases), os.path.getsize(fn) / 1e6))


def make_eh():
    """Eisenstein & Hu P(k, z), rs_drag, z_drag, growth factor / rate and sigma8 from the reference's engine
    (Cosmology(..., engine='eisenstein_hu'), flat LCDM without massive neutrinos) for Latin-hypercube cosmologies."""
    from cosmoprimo import Cosmology
    sys.path.insert(0, ROOT)
    from cosmoprimo_b200 import synthetic
    B = 8
    par = synthetic.lhs_cosmologies(B, seed=42)
    par = {name: np.concatenate([val, [synthetic.DESI_FIDUCIAL[name]]]) for name, val in par.items()}      # + DESI-like fiducial
    k = np.geomspace(1e-5, 1e2, 256)
    zs = np.array([0., 0.5, 3.])
    rs = np.array([1., 8., 20.])
    pk, derived, sigma8, sigma_rz, growth_rate_rz = [], [], [], [], []
    for i in range(B + 1):
        cosmo = Cosmology(h=par['h'][i], omega_b=par['omega_b'][i], omega_cdm=par['omega_cdm'][i], n_s=par['n_s'][i],
                          A_s=1e-10 * np.exp(par['logA'][i]), m_ncdm=None, engine='eisenstein_hu')
        fo, ba, th = cosmo.get_fourier(), cosmo.get_background(), cosmo.get_thermodynamics()
        interp = fo.pk_interpolator()
        pk.append(interp(k, z=zs).T)                                                                        # (nz, nk)
        derived.append([[th.rs_drag, th.z_drag, float(ba.growth_factor(z, znorm=0.))**2, float(ba.growth_rate(z))] for z in zs])
        sigma8.append(float(interp.sigma8_z(0.)))                     # integrate_sigma_r2(method='fftlog', nk=1024), interpolator.py:200, 285
        sigma_rz.append(interp.sigma_rz(rs, zs))                      # (nr, nz)
        growth_rate_rz.append(interp.growth_rate_rz(rs, zs))
        if i == 0:
            T_cmb, N_ur, k_pivot = float(cosmo['T_cmb']), float(cosmo['N_ur']), float(cosmo['k_pivot'])
            omega_r = float(ba.Omega0_r * cosmo['h']**2)
    arrays = dict(k=k, z=zs, pk=np.array(pk), derived=np.array(derived), sigma8=np.array(sigma8), r=rs, sigma_rz=np.array(sigma_rz), growth_rate_rz=np.array(growth_rate_rz), T_cmb=T_cmb, N_ur=N_ur, k_pivot=k_pivot,
                  omega_r=omega_r, **{'par_' + name: val for name, val in par.items()})
    fn = os.path.join(GOLDEN, 'eh_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({} cosmologies x {} redshifts x {} k, {:.2f} MB)'.format(fn, B + 1, zs.size, k.size, os.path.getsize(fn) / 1e6))


def make_interp2d():
    """PowerSpectrumInterpolator2D / CorrelationFunctionInterpolator2D of the reference (interpolator.py:608-987, 1219-1498;
    jax.py:213-277 = RectBivariateSpline) on a tabulated EH P(k, z): evaluation on grids and pairs, sigma_rz, sigma8_z,
    growth_rate_rz, to_1d, to_xi, to_pk, rescale_sigma8; with and without a growth_factor_sq callable."""
    from cosmoprimo.interpolator import PowerSpectrumInterpolator2D
    from cosmoprimo.jax import Interpolator2D
    sys.path.insert(0, ROOT)
    from cosmoprimo_b200 import synthetic
    arrays = {}
    k = np.geomspace(1e-4, 50., 300)
    z = np.array([0., 0.2, 0.5, 0.9, 1.4, 2., 3.])
    Om0, h = 0.3137721026737606, 0.6736
    D2 = synthetic.growth_factor(z, Om0, h)**2
    pk0 = synthetic.eh_pk(k)                                       # z = 0, DESI-like fiducial
    pk = pk0[:, None] * (D2 / D2[0]) * (1. + 0.05 * z * np.log10(k / 0.1)[:, None]**2 / 9.)     # mildly scale-dependent growth
    rng = np.random.default_rng(2)
    kq = np.concatenate([np.geomspace(2e-7, 90., 60), [1e-8, 200.]])
    zq = np.array([0., 0.1, 0.55, 1.7, 3., 3.5, -0.1])
    kp, zp = np.exp(rng.uniform(np.log(1e-6), np.log(80.), 50)), rng.uniform(0., 3., 50)
    r = np.array([1., 8., 20.])
    zs = np.array([0., 0.5, 1.1, 3.])
    arrays.update(k=k, z=z, pk=pk, kq=kq, zq=zq, kp=kp, zp=zp, r=r, zs=zs)
    # raw Interpolator2D (lin / log axes)
    i2 = Interpolator2D(k, z, pk, interp_x='log', interp_fun='log')
    arrays['i2_grid'] = i2(kq, zq)
    arrays['i2_pairs'] = i2(kp, zp, grid=False)
    i2 = Interpolator2D(np.log(k), z, np.log(pk), extrap=True)
    arrays['i2_lin_extrap'] = i2(np.log(kq), zq)
    interp = PowerSpectrumInterpolator2D(k, z, pk)
    arrays['p2_grid'] = interp(kq, zq)
    arrays['p2_pairs'] = interp(kp, zp, grid=False)
    arrays['p2_sigma_rz'] = interp.sigma_rz(r, zs)
    arrays['p2_sigma8_z'] = interp.sigma8_z(zs)
    arrays['p2_growth_rate_rz'] = interp.growth_rate_rz(r, zs)
    arrays['p2_to_1d'] = interp.to_1d(0.55)(kq)
    xi = interp.to_xi()
    sq = np.geomspace(xi.s[0] * 1.01, xi.s[-1] * 0.99, 40)
    arrays.update(sq=sq, xi_s=xi.s, p2_xi=xi(sq, zq[:5]), p2_xi_back=xi.to_pk(extrap_pk='lin')(np.geomspace(1e-3, 10., 30), zs))   # default extrap_pk='log' yields NaN: negative P(k) at the edges
    interp.rescale_sigma8(0.8)
    arrays['p2_rescaled'] = interp(kq[:20], zs)
    # single column + growth_factor_sq callable
    gf = lambda zz: np.interp(zz, z, D2 / D2[0])
    interp = PowerSpectrumInterpolator2D(k, 0., pk0, growth_factor_sq=gf)
    arrays['g_grid'] = interp(kq, zq)
    arrays['g_sigma_rz'] = interp.sigma_rz(r, zs)
    arrays['g_growth_rate_rz'] = interp.growth_rate_rz(r, zs)
    arrays['g_xi'] = interp.to_xi()(sq, zs)
    fn = os.path.join(GOLDEN, 'interp2d_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({:.2f} MB)'.format(fn, os.path.getsize(fn) / 1e6))


def make_filters():
    """The least-squares / peak-average BAO filters: reference outputs on EH spectra of one cosmology at four redshifts.
    The inputs the filters draw from the cosmology layer (rs_drag, no-wiggle spectra) are stored too, so that a failure can be attributed."""
    from cosmoprimo import Cosmology
    from cosmoprimo.cosmology import Fourier
    from cosmoprimo.bao_filter import PowerSpectrumBAOFilter, CorrelationFunctionBAOFilter
    par = dict(h=0.70, omega_b=0.0235, omega_cdm=0.125, n_s=0.95, A_s=2.2e-9)
    par_fid = dict(h=0.6736, omega_b=0.02237, omega_cdm=0.12, n_s=0.9649, A_s=2.083e-9)
    cosmo = Cosmology(m_ncdm=None, engine='eisenstein_hu', **par)
    cosmo_fid = Cosmology(m_ncdm=None, engine='eisenstein_hu', **par_fid)
    interp2d = cosmo.get_fourier().pk_interpolator()                     # from_callable: exact EH spectrum, growth factor applied per z
    ktab = np.geomspace(1e-5, 1e2, 1500)
    z = np.array([0., 0.5, 1., 2.])
    from cosmoprimo.interpolator import PowerSpectrumInterpolator1D
    interp = PowerSpectrumInterpolator1D(ktab, interp2d(ktab, z))        # tabulated, four columns
    arrays = dict(par=np.array([par[n] for n in ['h', 'omega_b', 'omega_cdm', 'n_s', 'A_s']]), par_fid=np.array([par_fid[n] for n in ['h', 'omega_b', 'omega_cdm', 'n_s', 'A_s']]),
                  rs_drag=cosmo.rs_drag, rs_drag_fid=cosmo_fid.rs_drag, extrap_kmin=interp.extrap_kmin, extrap_kmax=interp.extrap_kmax)
    f = PowerSpectrumBAOFilter(interp, engine='ehpoly', cosmo=cosmo)
    arrays.update(k=f.k, pk=f.pk, ehpoly_pknow=f.pknow, pk_nowiggle=Fourier(cosmo, engine='eisenstein_hu_nowiggle', set_engine=False).pk_interpolator()(f.k, z=0.))
    f = PowerSpectrumBAOFilter(interp, engine='ehpoly', cosmo=cosmo, cosmo_fid=cosmo_fid, krange=(2e-3, 0.8), rescale_krange=True)
    arrays.update(ehpoly2_pknow=f.pknow)
    f = PowerSpectrumBAOFilter(interp, engine='peakaverage', cosmo=cosmo, cosmo_fid=cosmo_fid)
    arrays.update(peakaverage_pknow=f.pknow, peakaverage_k_peaks0=f.k_peaks[0], peakaverage_k_peaks1=f.k_peaks[1], peakaverage_pad=np.array(f.pad_peaks))
    f1 = PowerSpectrumBAOFilter(PowerSpectrumInterpolator1D(ktab, interp2d(ktab, 0.5)), engine='peakaverage', cosmo=cosmo, cosmo_fid=cosmo_fid)
    arrays.update(peakaverage_pknow_1col=f1.pknow)
    xi = interp.to_xi()
    fx = CorrelationFunctionBAOFilter(xi, engine='kirkby2013', cosmo=cosmo)
    arrays.update(s=fx.s, xi=fx.xi, kirkby_xinow=fx.xinow, extrap_smin=xi.extrap_smin, extrap_smax=xi.extrap_smax)
    fx = CorrelationFunctionBAOFilter(xi, engine='kirkby2013', cosmo=cosmo, cosmo_fid=cosmo_fid, srange_left=(45., 80.), srange_right=(155., 195.))
    arrays.update(kirkby2_xinow=fx.xinow)
    fn = os.path.join(GOLDEN, 'filters_golden.npz')
    np.savez_compressed(fn, **arrays)
    print('wrote {} ({:.2f} MB)'.format(fn, os.path.getsize(fn) / 1e6))


if __name__ == '__main__':
    which = sys.argv[1:] or ['fftlog', 'spline', 'wallish', 'eh', 'interp2d', 'filters']
    print('reference: cosmoprimo {} from {}; numpy {}'.format(cosmoprimo.__version__, os.path.dirname(cosmoprimo.__file__), np.__version__))
    for name in which:
        fn = globals().get('make_' + name, None)
        if fn is None:
            print('skipping {} (no generator yet)'.format(name))
            continue
        fn()
