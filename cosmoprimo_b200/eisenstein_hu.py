"""
On-device Eisenstein & Hu (1998) linear matter power spectra for batches of flat LCDM cosmologies without massive
neutrinos: what the reference computes, one cosmology at a time, with
``Cosmology(..., m_ncdm=None, engine='eisenstein_hu').get_fourier().pk_interpolator()(k, z)``
(``cosmoprimo/eisenstein_hu.py`` — coefficients :34-92, transfer function :241-283, primordial spectrum :189-214, P(k)
:321-324, growth factor with ``znorm=0`` / growth rate :115-153 on the background of ``cosmoprimo/cosmology.py:1675-1760``).

The kernel (``csrc/cpf_eh.cu`` behind ``cpf_eh_pk``) writes rows in the (rows, nk) layout FFTLog reads, so a sweep over
cosmologies x redshifts runs generator -> FFTLog without any host-to-device copy of spectra.
"""

import numpy as np

from . import _lib
from . import _buffers as _buf

T_CMB = 2.7255
N_UR = 3.044
K_PIVOT = 0.05     # 1/Mpc


def omega_radiation(T_cmb=T_CMB, N_ur=N_UR):
    """``Omega0_r h^2``: photons + ``N_ur`` massless neutrinos, with the reference's constants (cosmology.py:355-367,
    constants.py:12-14; scipy.constants for c, G, sigma_SB and the parsec)."""
    from scipy import constants
    megaparsec_over_m = 1e6 * constants.parsec
    rho_crit = 3.0 * (100. * 1e3 / megaparsec_over_m)**2 / (8 * constants.pi * constants.gravitational_constant)   # h^2 kg/m^3
    T_ur = T_cmb * (4. / 11.)**(1. / 3.)
    rho = (T_cmb**4 + N_ur * 7. / 8. * T_ur**4) * 4. / constants.c**3 * constants.Stefan_Boltzmann
    return rho / rho_crit


class EisensteinHu(object):
    """
    Batch of B cosmologies; parameters are scalars or (B,) arrays (broadcast against each other).

    Parameters
    ----------
    h, omega_b, omega_cdm, n_s : array_like
    A_s or logA : array_like
        Scalar amplitude, or ``ln(1e10 A_s)`` (the reference's ``logA``, cosmology.py:949-950).
    T_cmb, N_ur, k_pivot : float
        Reference defaults (constants.py:17-19, cosmology.py defaults); ``k_pivot`` in 1/Mpc.
    device : int, default=None
        CUDA device.
    """

    def __init__(self, h, omega_b, omega_cdm, n_s, A_s=None, logA=None, T_cmb=T_CMB, N_ur=N_UR, k_pivot=K_PIVOT, device=None):
        if (A_s is None) == (logA is None):
            raise ValueError('provide either A_s or logA')
        if A_s is None:
            A_s = 1e-10 * np.exp(np.asarray(logA, dtype='f8'))
        cols = np.broadcast_arrays(*[np.asarray(v, dtype='f8') for v in (h, omega_b, omega_cdm, n_s, A_s)])
        self.shape = cols[0].shape
        self.params = np.ascontiguousarray(np.stack([c.ravel() for c in cols], axis=-1))      # (B, 5)
        self.T_cmb, self.N_ur, self.k_pivot = float(T_cmb), float(N_ur), float(k_pivot)
        self.omega_r = omega_radiation(self.T_cmb, self.N_ur)
        self.device = device

    @property
    def size(self):
        return self.params.shape[0]

    def _run(self, k, z, kaiser, on_device, want_pk=True):
        lib = _lib.load()
        _lib.require_device()
        B = self.size
        zz, nz, zshape = None, 1, ()
        if z is not None:
            zz = np.asarray(z, dtype='f8')
            if zz.ndim == 2:                                   # (B, nz): a redshift grid per cosmology
                if zz.shape[0] not in (1, B):
                    raise ValueError('z must have shape (), ({0},), (1, nz) or ({0}, nz)'.format(B))
                nz = zz.shape[1]
                zshape = (nz,)
                zz = np.broadcast_to(zz, (B, nz))
            else:                                              # scalar or (B,): one redshift per cosmology
                zz = np.broadcast_to(zz, (B,))
            zz = np.array(zz, dtype='f8', order='C')
        dev = self.device if self.device is not None else _buf.default_device()
        pshape = (3,) if kaiser else ()
        kk = np.ascontiguousarray(k, dtype='f8').ravel() if want_pk else np.ones(1)
        nk = kk.size
        oshape = (B,) + zshape + pshape + (nk,)
        dshape = (B,) + zshape + (4,)
        if on_device:
            torch = _buf._torch()
            tdev = torch.device('cuda', dev)
            params = torch.as_tensor(self.params, device=tdev)
            kd = torch.as_tensor(kk, device=tdev)
            zd = torch.as_tensor(zz, device=tdev) if zz is not None else None
            out = torch.empty(oshape, dtype=torch.float64, device=tdev)
            derived = torch.empty(dshape, dtype=torch.float64, device=tdev)
            rc = lib.cpf_eh_pk(params.data_ptr(), zd.data_ptr() if zd is not None else None, B, nz, kd.data_ptr(), nk, self.T_cmb,
                               self.omega_r, self.k_pivot, int(kaiser), out.data_ptr(), derived.data_ptr(), 1, dev, _buf.current_stream(dev))
        else:
            out = np.empty(oshape, dtype='f8')
            derived = np.empty(dshape, dtype='f8')
            rc = lib.cpf_eh_pk(self.params.ctypes.data, zz.ctypes.data if zz is not None else None, B, nz, kk.ctypes.data, nk, self.T_cmb,
                               self.omega_r, self.k_pivot, int(kaiser), out.ctypes.data, derived.ctypes.data, 0, dev, None)
        _lib.check(rc)
        return out, derived

    def pk(self, k, z=None, kaiser=False, on_device=True):
        """
        Linear P(k, z) in (Mpc/h)^3 on ``k`` [h/Mpc].  ``z``: scalar or (B,) -- one redshift per cosmology, result
        (B, nk) -- or (B, nz) / (1, nz) -- a redshift grid per cosmology, result (B, nz, nk); the transfer function is
        evaluated once per cosmology.  ``kaiser``: the Kaiser multipoles ell = 0, 2, 4 with f = growth_rate(z) instead,
        (..., 3, nk).  Returns a torch CUDA tensor (``on_device``) or a numpy array.
        """
        return self._run(k, z, kaiser, on_device)[0]

    def derived(self, z=None, on_device=False):
        """(..., 4): rs_drag [Mpc/h], z_drag, growth_factor(z, znorm=0)**2, growth_rate(z); leading shape as :meth:`pk`."""
        return self._run(None, z, False, on_device, want_pk=False)[1]


class EHCosmology(object):
    """
    The handful of cosmology-dependent numbers the polynomial / peak-average BAO filters need (``cosmoprimo/bao_filter.py`` uses
    ``cosmo.rs_drag`` and ``Fourier(cosmo, engine='eisenstein_hu_nowiggle').pk_interpolator()(k, z=0)``): one flat LCDM cosmology without
    massive neutrinos, evaluated with the Eisenstein & Hu formulae exactly as the reference's engines do
    (``cosmoprimo/eisenstein_hu.py:34-66`` for ``rs_drag``, ``cosmoprimo/eisenstein_hu_nowiggle.py:17-51`` for the no-wiggle transfer function,
    ``cosmoprimo/eisenstein_hu.py:189-214, 321-324`` for the primordial spectrum and the potential / curvature factors).  Host numpy code: one
    cosmology per filter, nothing batched.  Reference ``Cosmology`` objects are accepted by the filters as well (duck typing).
    """

    def __init__(self, h, omega_b, omega_cdm, n_s, A_s=None, logA=None, T_cmb=T_CMB, k_pivot=K_PIVOT):
        if (A_s is None) == (logA is None):
            raise ValueError('provide either A_s or logA')
        self.h, self.omega_b, self.omega_cdm, self.n_s = float(h), float(omega_b), float(omega_cdm), float(n_s)
        self.A_s = float(A_s) if A_s is not None else 1e-10 * float(np.exp(logA))
        self.T_cmb, self.k_pivot = float(T_cmb), float(k_pivot)
        # ref eisenstein_hu.py:34-66
        self.omega_m = self.omega_cdm + self.omega_b
        self.frac_b = self.omega_b / self.omega_m
        self.theta_cmb = self.T_cmb / 2.7
        self.z_eq = 2.5e4 * self.omega_m * self.theta_cmb**(-4) - 1.
        self.k_eq = 0.0746 * self.omega_m * self.theta_cmb**(-2)
        b1 = 0.313 * self.omega_m**(-0.419) * (1 + 0.607 * self.omega_m**0.674)
        b2 = 0.238 * self.omega_m**0.223
        self.z_drag = 1345 * self.omega_m**0.251 / (1. + 0.659 * self.omega_m**0.828) * (1. + b1 * self.omega_b**b2)
        r_drag = 31.5 * self.omega_b * self.theta_cmb**(-4) * (1000. / (1 + self.z_drag))
        r_eq = 31.5 * self.omega_b * self.theta_cmb**(-4) * (1000. / (1 + self.z_eq))
        self._rs_drag_mpc = 2. / (3. * self.k_eq) * np.sqrt(6. / r_eq) * np.log((np.sqrt(1 + r_drag) + np.sqrt(r_drag + r_eq)) / (1 + np.sqrt(r_eq)))
        # ref eisenstein_hu_nowiggle.py:20-23
        self.alpha_gamma = 1. - 0.328 * np.log(431. * self.omega_m) * self.frac_b + 0.38 * np.log(22.3 * self.omega_m) * self.frac_b**2

    @property
    def rs_drag(self):
        """Sound horizon at the drag epoch in Mpc/h, as ``Cosmology.rs_drag`` of the reference."""
        return self._rs_drag_mpc * self.h

    def _pk_from_transfer(self, k, transfer):
        # ref eisenstein_hu.py:189-214 (primordial, k_pivot in 1/Mpc) and :321-324; growth factor at z = 0 with the reference's `znorm=0` convention
        from . import synthetic
        k = np.asarray(k, dtype='f8')
        Omega0_m = self.omega_m / self.h**2
        potential_to_density = (3. * Omega0_m * 100**2 / (2. * synthetic.C_KMS**2 * k**2))**(-2)
        curvature_to_potential = 9. / 25. * 2. * np.pi**2 / k**3 / self.h**3
        primordial = self.h**3 * self.A_s * (k / (self.k_pivot / self.h))**(self.n_s - 1.)
        return transfer**2 * potential_to_density * curvature_to_potential * primordial * synthetic.growth_factor(0., Omega0_m, self.h)**2

    def transfer_nowiggle(self, k):
        """No-wiggle transfer function on ``k`` [h/Mpc] (ref eisenstein_hu_nowiggle.py:35-51)."""
        kk = np.asarray(k, dtype='f8') * self.h
        ks = kk * self._rs_drag_mpc
        gamma_eff = self.omega_m * (self.alpha_gamma + (1 - self.alpha_gamma) / (1 + (0.43 * ks)**4))
        q = kk * self.theta_cmb**2 / gamma_eff
        L0 = np.log(2 * np.e + 1.8 * q)
        C0 = 14.2 + 731.0 / (1 + 62.5 * q)
        return L0 / (L0 + C0 * q**2)

    def pk_nowiggle(self, k):
        """Linear no-wiggle P(k, z=0) [(Mpc/h)^3]."""
        return self._pk_from_transfer(k, self.transfer_nowiggle(k))

    def pk_lin(self, k):
        """Linear P(k, z=0) with baryon wiggles (Eisenstein & Hu 1998 full fitting formula)."""
        from . import synthetic
        return synthetic.eh_pk(k, dict(h=self.h, omega_b=self.omega_b, omega_cdm=self.omega_cdm, n_s=self.n_s, logA=np.log(1e10 * self.A_s)), z=0.)
