"""
ORACLE (test infrastructure, not product code): CPU restatement of cosmoprimo's FFTLog path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this module.  The product package ``cosmoprimo_b200`` never does.

Every function cites the reference lines (relative to /root/reference/cosmoprimo/) that it restates.
The arithmetic that the reference delegates to third-party wheels is delegated to the *same* wheels here
(``numpy.fft`` = pocketfft, ``scipy.special.loggamma``; cosmoprimo's pyproject.toml:12 pins neither), so on one
machine this oracle is bit-identical to the reference numpy engine.  Parity is pinned by
``tests/test_oracle_golden.py`` against vectors produced by the reference itself (``tools/make_golden.py``) and by
the reference's own analytic known-answer test (tests/test_fftlog.py:56-89).

A second, independent restatement of the execute step (``execute_direct``: O(N^2) DFT sums in long double, no FFT
library at all) is provided to pin the FFT semantics (sign, conj, irfft's dropped imaginary parts).
"""

import numpy as np
from scipy.special import loggamma, gamma


# ----------------------------------------------------------------------------------------------------------------
# Mellin-transformed kernels U(z)                                            fftlog.py:666-766
# ----------------------------------------------------------------------------------------------------------------

def mellin_kernel(name, z, **p):
    """U_K(z) = int_0^inf t^(z-1) K(t) dt for the kernels of fftlog.py:688-766."""
    z = np.asarray(z)
    ln2 = np.log(2)
    if name == 'bessel_j':              # fftlog.py:695
        nu = p['nu']
        return np.exp(ln2 * (z - 1) + loggamma(0.5 * (nu + z)) - loggamma(0.5 * (2 + nu - z)))
    if name == 'spherical_bessel_j':    # fftlog.py:705
        nu = p['nu']
        return np.exp(ln2 * (z - 1.5) + loggamma(0.5 * (nu + z)) - loggamma(0.5 * (3 + nu - z)))
    if name == 'tophat':                # fftlog.py:726
        d = p.get('ndim', 1)
        return np.exp(ln2 * (z - 1) + loggamma(1 + 0.5 * d) + loggamma(0.5 * z) - loggamma(0.5 * (2 + d - z)))
    if name == 'tophat_sq':             # fftlog.py:739-746
        d = p.get('ndim', 1)
        if d == 1:
            return -0.25 * np.sqrt(np.pi) * np.exp(loggamma(0.5 * (z - 2)) - loggamma(0.5 * (3 - z)))
        if d == 3:
            return 2.25 * np.sqrt(np.pi) * (z - 2) / (z - 6) * np.exp(loggamma(0.5 * (z - 4)) - loggamma(0.5 * (5 - z)))
        return np.exp(ln2 * (d - 1) + 2 * loggamma(1 + 0.5 * d) + loggamma(0.5 * (1 + d - z)) + loggamma(0.5 * z)
                      - loggamma(1 + d - 0.5 * z) - loggamma(0.5 * (2 + d - z))) / np.sqrt(np.pi)
    if name == 'gaussian':              # fftlog.py:756
        return 2**(0.5 * z - 1) * gamma(0.5 * z)
    if name == 'gaussian_sq':           # fftlog.py:766
        return 0.5 * gamma(0.5 * z)
    raise ValueError('unknown kernel {}'.format(name))


# ----------------------------------------------------------------------------------------------------------------
# pad                                                                         fftlog.py:436-505
# ----------------------------------------------------------------------------------------------------------------

def pad_last(a, left, right, extrap=0):
    """Pad the last axis of ``a`` by (left, right) samples; ``extrap`` is a scalar, 'edge', 'log' or a (left, right) pair."""
    a = np.asarray(a)
    try:
        el, er = extrap
        if isinstance(extrap, str): raise TypeError
    except (TypeError, ValueError):
        el = er = extrap
    lead = a.shape[:-1]

    def side(mode, count, is_left):
        if isinstance(mode, str) and mode == 'edge':                     # :483-485, :494-496
            return np.repeat(a[..., :1] if is_left else a[..., -1:], count, axis=-1)
        if isinstance(mode, str) and mode == 'log':
            if is_left:                                                  # :486-490
                j = np.arange(-count, 0)
                return a[..., :1] * (a[..., 1:2] / a[..., :1]) ** j
            j = np.arange(1, count + 1)                                  # :497-501
            return a[..., -1:] / (a[..., -2:-1] / a[..., -1:]) ** j
        return np.full(lead + (count,), mode)                            # :492, :503

    return np.concatenate([side(el, left, True), a, side(er, right, False)], axis=-1)


# ----------------------------------------------------------------------------------------------------------------
# plan                                                                        fftlog.py:49-117, 144-184
# ----------------------------------------------------------------------------------------------------------------

def make_plan(x, kernels, q=0., minfolds=2, lowring=True, xy=1.):
    """
    Build the FFTLog tables.  ``kernels``: one (name, params) tuple, or a list of them (=> ``inparallel``).
    Returns a dict with the reference's attribute names.
    """
    inparallel = isinstance(kernels, list)
    if not inparallel: kernels = [kernels]
    P = len(kernels)
    qs = [q] * P if np.ndim(q) == 0 else list(q)
    xys = [xy] * P if np.ndim(xy) == 0 else list(xy)
    x = np.asarray(x, dtype='f8')
    x = np.tile(x[None, :], (P, 1)) if x.ndim == 1 else x                # :101-105
    n = x.shape[-1]
    delta = np.log(x[:, -1] / x[:, 0]) / (n - 1)                         # :147
    N = 2 ** ((n * minfolds - 1).bit_length())                           # :149-150
    npad = N - n
    in_left, in_right = npad // 2, npad - npad // 2                      # :152
    out_left, out_right = npad - npad // 2, npad // 2                    # :153
    U = [lambda z, k=k: mellin_kernel(k[0], z, **k[1]) for k in kernels]
    if lowring:                                                          # :162
        lnxy = np.array([d / np.pi * np.angle(u(qq + 1j * np.pi / d)) for u, d, qq in zip(U, delta, qs)], dtype='f8')
    else:                                                                # :164
        lnxy = np.log(xys) + delta
    y = np.exp(lnxy - delta)[:, None] / x[:, ::-1]                       # :166
    m = np.arange(0, N // 2 + 1)                                         # :168
    px = pad_last(x, in_left, in_right, 'log')                           # :170
    py = pad_last(y, out_left, out_right, 'log')                         # :171
    pre = np.array([px[p] ** (-qs[p]) for p in range(P)])                # :174
    post = np.array([py[p] ** (-qs[p]) for p in range(P)])               # :175
    u = np.array([U[p](qs[p] + 2j * np.pi / N / delta[p] * m) * np.exp(-2j * np.pi * lnxy[p] / N / delta[p] * m)
                  for p in range(P)])                                    # :179-180
    return dict(inparallel=inparallel, x=x, y=y, delta=delta, lnxy=lnxy, n=n, N=N, P=P,
                in_left=in_left, in_right=in_right, out_left=out_left, out_right=out_right,
                padded_x=px, padded_y=py, padded_u=u, padded_prefactor=pre, padded_postfactor=post)


def _ells(ell):
    return np.atleast_1d(ell)


def _sph_kernels(ell):
    if np.ndim(ell) == 0: return ('spherical_bessel_j', dict(nu=ell))
    return [('spherical_bessel_j', dict(nu=l)) for l in ell]


def plan_hankel(x, nu=0, **kw):                                          # fftlog.py:258-280
    k = ('bessel_j', dict(nu=nu)) if np.ndim(nu) == 0 else [('bessel_j', dict(nu=v)) for v in nu]
    pl = make_plan(x, k, **kw)
    pl['padded_prefactor'] = pl['padded_prefactor'] * (pl['padded_x'] ** 2)
    return pl


def plan_power_to_correlation(k, ell=0, q=0, complex=False, **kw):       # fftlog.py:292-330
    pl = make_plan(k, _sph_kernels(ell), q=1.5 + np.asarray(q) if np.ndim(q) else 1.5 + q, **kw)
    pl['padded_prefactor'] = pl['padded_prefactor'] * (pl['padded_x'] ** 3 / (2 * np.pi) ** 1.5)
    phase = (-1j) ** _ells(ell) if complex else (-1) ** (_ells(ell) // 2)
    pl['padded_postfactor'] = pl['padded_postfactor'] * phase[:, None]
    return pl


def plan_correlation_to_power(s, ell=0, q=0, complex=False, **kw):       # fftlog.py:342-377
    pl = make_plan(s, _sph_kernels(ell), q=1.5 + np.asarray(q) if np.ndim(q) else 1.5 + q, **kw)
    pl['padded_prefactor'] = pl['padded_prefactor'] * (pl['padded_x'] ** 3 * (2 * np.pi) ** 1.5)
    phase = (1j) ** _ells(ell) if complex else (-1) ** (_ells(ell) // 2)
    pl['padded_postfactor'] = pl['padded_postfactor'] * phase[:, None]
    return pl


def plan_tophat_variance(k, q=0, **kw):                                  # fftlog.py:387-405
    pl = make_plan(k, ('tophat_sq', dict(ndim=3)), q=1.5 + q, **kw)
    pl['padded_prefactor'] = pl['padded_prefactor'] * (pl['padded_x'] ** 3 / (2 * np.pi ** 2))
    return pl


def plan_gaussian_variance(k, q=0, **kw):                                # fftlog.py:415-433
    pl = make_plan(k, ('gaussian_sq', {}), q=1.5 + q, **kw)
    pl['padded_prefactor'] = pl['padded_prefactor'] * (pl['padded_x'] ** 3 / (2 * np.pi ** 2))
    return pl


def invert_plan(pl):                                                     # fftlog.py:243-248
    """In-place inverse, with the reference's quirk of storing the unpadded grids in padded_x/padded_y (:246)."""
    pl['x'], pl['y'] = pl['y'], pl['x']
    pl['padded_x'], pl['padded_y'] = pl['y'], pl['x']
    pl['padded_prefactor'], pl['padded_postfactor'] = 1 / pl['padded_postfactor'], 1 / pl['padded_prefactor']
    pl['padded_u'] = 1 / pl['padded_u'].conj()
    return pl


# ----------------------------------------------------------------------------------------------------------------
# execute                                                                     fftlog.py:198-241, 538-544
# ----------------------------------------------------------------------------------------------------------------

def execute(pl, fun, extrap=0, keep_padding=False):
    """(y, G) exactly as FFTlog.__call__ with the numpy engine."""
    fun = np.asarray(fun)
    a = pad_last(fun, pl['in_left'], pl['in_right'], extrap) * pl['padded_prefactor']          # :230-231
    A = np.fft.rfft(a, axis=-1)                                                                # :540
    g = np.fft.irfft((A * pl['padded_u']).conj(), n=pl['N'], axis=-1)                          # :544
    G = g * pl['padded_postfactor']
    if not keep_padding:                                                                       # :233-237
        y = pl['y']
        G = G[..., pl['out_left']:pl['out_left'] + pl['n']]
    else:
        y = pl['padded_y']
    if not pl['inparallel']:                                                                   # :238-240
        y = y[0]
        G = np.reshape(G, fun.shape if not keep_padding else fun.shape[:-1] + (pl['N'],))
    return y, G


def execute_direct(pl, fun, extrap=0):
    """
    Library-free restatement of the execute step for ONE input row and plan row 0, O(N^2) in long double
    (SURVEY Appendix A.6): A[m] = sum_j a[j] e^{-2 pi i jm/N}; C = conj(A u);
    g[j] = (Re C[0] + (-1)^j Re C[N/2] + 2 Re sum_{m=1}^{N/2-1} C[m] e^{+2 pi i jm/N}) / N.
    Returns the cropped G.  Small N only.
    """
    N, n = pl['N'], pl['n']
    ld = np.longdouble
    a = (pad_last(np.asarray(fun, dtype='f8'), pl['in_left'], pl['in_right'], extrap) * pl['padded_prefactor'][0]).astype(ld)
    j = np.arange(N)
    m = np.arange(N // 2 + 1)
    ang = (2 * ld(np.pi) / N) * ((j[:, None] * m[None, :]) % N).astype(ld)     # (N, N/2+1)
    c, s = np.cos(ang), np.sin(ang)
    Are = (a[:, None] * c).sum(axis=0)
    Aim = -(a[:, None] * s).sum(axis=0)
    ure, uim = pl['padded_u'][0].real.astype(ld), pl['padded_u'][0].imag.astype(ld)
    Cre = Are * ure - Aim * uim
    Cim = -(Are * uim + Aim * ure)
    w = np.full(N // 2 + 1, 2, dtype=ld)
    w[0] = w[-1] = 1
    Cim_eff = Cim.copy()
    Cim_eff[0] = Cim_eff[-1] = 0     # irfft discards Im at DC and Nyquist
    g = ((w * Cre)[None, :] * c - (w * Cim_eff)[None, :] * s).sum(axis=1) / N
    G = g * pl['padded_postfactor'][0]
    return np.asarray(G[pl['out_left']:pl['out_left'] + n])


# ----------------------------------------------------------------------------------------------------------------
# parity metric                                                               SURVEY.md §8(d)
# ----------------------------------------------------------------------------------------------------------------

def scale_aware_error(G, G_ref, post):
    """
    max_j |G - G_ref| |w| / max_j |G_ref w| per row, with w = 1/post (compare in the biased space G y^q where the
    FFT's rounding error is uniform).  ``post`` must broadcast against G.
    """
    w = 1. / np.abs(post)
    num = np.max(np.abs(G - G_ref) * w, axis=-1)
    den = np.max(np.abs(G_ref) * w, axis=-1)
    return num / den
