"""
Synthetic inputs for tests and benchmarks: Eisenstein & Hu (1998) linear matter power spectra for Latin-hypercube
cosmologies.  Host numpy code — an input generator, not part of the accelerated path.

Formulas: Eisenstein & Hu 1998 (ApJ 496, 605) eqs. 2-24 as used by ``cosmoprimo/eisenstein_hu.py:34-92, 241-283``
(with its z_drag normalisation 1345, :53), primordial spectrum and potential/curvature factors of ``:189-215,
321-324``, growth factor of Carroll, Press & Turner 1992 (``:115-140``).  Flat LCDM, massless neutrinos.
"""

import numpy as np

T_CMB = 2.7255
C_KMS = 299792.458
K_PIVOT = 0.05     # 1/Mpc

# ranges of the reference's emulator training set (cosmoprimo/emulators/train/train_classy.py:63-64)
PARAM_NAMES = ('logA', 'n_s', 'h', 'omega_b', 'omega_cdm')
PARAM_LOWER = np.array([2.5, 0.88, 0.5, 0.019, 0.08])
PARAM_UPPER = np.array([3.5, 1.06, 0.9, 0.026, 0.2])
DESI_FIDUCIAL = dict(logA=np.log(2.083e-9 * 1e10), n_s=0.9649, h=0.6736, omega_b=0.02237, omega_cdm=0.12)


def lhs_cosmologies(size, seed=42):
    """``size`` cosmologies drawn like the reference's ``QMCSampler(engine='lhs')`` (emulators/tools/samples.py:701-713)."""
    from scipy.stats import qmc
    sample = qmc.LatinHypercube(d=len(PARAM_NAMES), seed=seed).random(n=size)
    sample = qmc.scale(sample, PARAM_LOWER, PARAM_UPPER)
    return {name: sample[:, i] for i, name in enumerate(PARAM_NAMES)}


def eh_transfer(k, h, omega_b, omega_m):
    """EH98 transfer function with baryon wiggles; ``k`` [h/Mpc] of shape (nk,), parameters of shape (B, 1)."""
    theta = T_CMB / 2.7
    fb = omega_b / omega_m
    z_eq = 2.5e4 * omega_m * theta**-4 - 1.
    k_eq = 0.0746 * omega_m * theta**-2
    b1 = 0.313 * omega_m**-0.419 * (1 + 0.607 * omega_m**0.674)
    b2 = 0.238 * omega_m**0.223
    z_drag = 1345 * omega_m**0.251 / (1. + 0.659 * omega_m**0.828) * (1. + b1 * omega_b**b2)
    r_drag = 31.5 * omega_b * theta**-4 * (1000. / (1 + z_drag))
    r_eq = 31.5 * omega_b * theta**-4 * (1000. / (1 + z_eq))
    rs = 2. / (3. * k_eq) * np.sqrt(6. / r_eq) * np.log((np.sqrt(1 + r_drag) + np.sqrt(r_drag + r_eq)) / (1 + np.sqrt(r_eq)))
    k_silk = 1.6 * omega_b**0.52 * omega_m**0.73 * (1 + (10.4 * omega_m)**-0.95)
    a1 = (46.9 * omega_m)**0.670 * (1 + (32.1 * omega_m)**-0.532)
    a2 = (12.0 * omega_m)**0.424 * (1 + (45.0 * omega_m)**-0.582)
    alpha_c = a1**-fb * a2**(-fb**3)
    bb1 = 0.944 / (1 + (458 * omega_m)**-0.708)
    bb2 = 0.395 * omega_m**-0.0266
    beta_c = 1. / (1 + bb1 * ((1 - fb)**bb2) - 1)
    yd = (1 + z_eq) / (1 + z_drag)
    G = yd * (-6. * np.sqrt(1 + yd) + (2. + 3. * yd) * np.log((np.sqrt(1 + yd) + 1) / (np.sqrt(1 + yd) - 1)))
    alpha_b = 2.07 * k_eq * rs * (1 + r_drag)**-0.75 * G
    beta_node = 8.41 * omega_m**0.435
    beta_b = 0.5 + fb + (3. - 2. * fb) * np.sqrt((17.2 * omega_m)**2 + 1)

    kk = k * h                                   # 1/Mpc
    q = kk / (13.41 * k_eq)
    ks = kk * rs
    ln_beta = np.log(np.e + 1.8 * beta_c * q)
    ln_nobeta = np.log(np.e + 1.8 * q)
    C_alpha = 14.2 / alpha_c + 386. / (1 + 69.9 * q**1.08)
    C_noalpha = 14.2 + 386. / (1 + 69.9 * q**1.08)
    f = 1. / (1. + (ks / 5.4)**4)
    T0 = lambda a, b: a / (a + b * q**2)
    Tc = f * T0(ln_beta, C_noalpha) + (1 - f) * T0(ln_beta, C_alpha)
    s_tilde = rs * (1 + (beta_node / ks)**3)**(-1. / 3.)
    Tb = np.sinc(kk * s_tilde / np.pi) * (T0(ln_nobeta, C_noalpha) / (1 + (ks / 5.2)**2)
                                           + alpha_b / (1 + (beta_b / ks)**3) * np.exp(-(kk / k_silk)**1.4))
    return fb * Tb + (1 - fb) * Tc


def omega_radiation():
    """Omega0_r h^2 of the reference's default background (photons at T_cmb + N_ur = 3.044 massless neutrinos)."""
    from .eisenstein_hu import omega_radiation as _omega_r
    return _omega_r(T_CMB, 3.044)


def background(z, Omega0_m, h):
    """Omega_m(z), Omega_de(z) of flat LCDM with radiation, as the reference's background (cosmology.py:1675-1760)."""
    Omega0_r = omega_radiation() / h**2
    Omega0_de = 1. - Omega0_m - Omega0_r
    crit = Omega0_m + Omega0_r * (1 + z) + Omega0_de / (1 + z)**3
    return Omega0_m / crit, Omega0_de / (1 + z)**3 / crit


def growth_factor(z, Omega0_m, h=None):
    """Carroll-Press-Turner growth factor, flat LCDM, normalised to 1/(1+z) in matter domination — the ``znorm=0``
    convention the reference's Fourier.pk_interpolator applies (eisenstein_hu.py:115-140, 319).  With ``h`` the
    background includes radiation exactly as the reference's does; without, matter + Lambda only."""
    if h is None:
        E2 = Omega0_m * (1 + z)**3 + 1. - Omega0_m
        Om, Ode = Omega0_m * (1 + z)**3 / E2, (1. - Omega0_m) / E2
    else:
        Om, Ode = background(z, Omega0_m, h)
    return 1. / (1 + z) * 5 * Om / 2. / (Om**(4. / 7.) - Ode + (1. + Om / 2.) * (1 + Ode / 70.))


def growth_rate(z, Omega0_m, h):
    """f(z) = Omega_m(z)^0.55 (eisenstein_hu.py:141-153 for w = -1)."""
    return background(z, Omega0_m, h)[0]**0.55


def eh_pk(k, params=None, z=0.):
    """
    Linear P(k) [(Mpc/h)^3] on ``k`` [h/Mpc] for a dict of parameter arrays (default: DESI fiducial); returns
    (B, nk) for B cosmologies (or (nk,) for scalars).  ``z`` scalar or (B,) array.
    """
    params = dict(DESI_FIDUCIAL if params is None else params)
    scalar = np.ndim(params['h']) == 0
    p = {name: np.atleast_1d(np.asarray(params[name], dtype='f8'))[:, None] for name in PARAM_NAMES}
    k = np.asarray(k, dtype='f8')
    h = p['h']
    omega_m = p['omega_b'] + p['omega_cdm']
    Omega0_m = omega_m / h**2
    A_s = 1e-10 * np.exp(p['logA'])
    T = eh_transfer(k, h, p['omega_b'], omega_m)
    potential_to_density = (3. * Omega0_m * 100**2 / (2. * C_KMS**2 * k**2))**-2
    curvature_to_potential = 9. / 25. * 2. * np.pi**2 / k**3 / h**3
    primordial = h**3 * A_s * (k / (K_PIVOT / h))**(p['n_s'] - 1.)
    pk = T**2 * potential_to_density * curvature_to_potential * primordial
    D = growth_factor(np.atleast_1d(np.asarray(z, dtype='f8'))[:, None] if np.ndim(z) else z, Omega0_m, h)
    pk = pk * D**2
    return pk[0] if scalar else pk


def kaiser_multipoles(pk, f):
    """(B, nk) -> (B, 3, nk): ell = 0, 2, 4 Kaiser multipoles for growth rate ``f``."""
    f = np.asarray(f, dtype='f8')
    if f.ndim: f = f[:, None]
    return np.stack([(1 + 2 * f / 3 + f**2 / 5) * pk, (4 * f / 3 + 4 * f**2 / 7) * pk, (8 * f**2 / 35) * pk], axis=-2)
