"""
The callers either side of the FFTLog path, GPU-backed: 1-D power-spectrum / correlation-function interpolators with
the reference's call signatures (``cosmoprimo/interpolator.py``, cited as ``ref:LINE``).  Only what sits on the hot
path is here: log-log extrapolation padding (``_pad_log``, ref:42-87), spline evaluation (ref:495-521),
``sigma_r`` with ``method='fftlog'`` (ref:200-292, branch 285-289), ``to_xi`` / ``to_pk`` (ref:584-605,
1194-1215) and their 2-D (k, z) counterparts (ref:608-987, 1219-1498: bicubic table, ``sigma_rz``, ``sigma8_z``,
``growth_rate_rz``, ``to_1d``, ``to_xi`` / ``to_pk``).  ``from_callable`` constructors, quadrature methods and pytree
plumbing stay in the reference.

numpy tables in -> numpy results; CUDA tables in -> torch results (the spline lives on the device either way).
"""

import numpy as np

from . import _buffers as _buf
from .interp import Interpolator1D, Interpolator2D, _bcast_dtype, spline_eval_rows
from .fftlog import PowerToCorrelation, CorrelationToPower, TophatVariance

_default_extrap_kmin = 1e-7
_default_extrap_kmax = 1e2


def _xp(a):
    """numpy, or torch for CUDA arrays."""
    return _buf._torch() if _buf.is_device_array(a) else np


def _pad_log(k, pk, extrap_kmin=_default_extrap_kmin, extrap_kmax=_default_extrap_kmax):
    """
    log10(k), log10(pk) with two extra knots on each side that continue ``pk`` as a power law down to ``extrap_kmin``
    and up to ``extrap_kmax`` (ref:42-87): slopes from the two edge samples, knots at the range end and 90 % of the way
    to it.  ``k`` is a host array (nk,), ``pk`` (nk, ...) numpy or torch.
    """
    xp = _xp(pk)
    logk, pad_lo, pad_hi = _pad_log_knots(k, extrap_kmin=extrap_kmin, extrap_kmax=extrap_kmax)
    logpk = xp.log10(pk)
    slope_hi = (logpk[-1] - logpk[-2]) / (logk[-1] - logk[-2])
    slope_lo = (logpk[1] - logpk[0]) / (logk[1] - logk[0])
    rows_hi = [logpk[-1] + slope_hi * (x - logk[-1]) for x in pad_hi]
    rows_lo = [logpk[0] + slope_lo * (x - logk[0]) for x in pad_lo]
    stack = xp.stack if xp is not np else np.stack
    cat = xp.cat if xp is not np else np.concatenate
    logpk = cat([stack(rows_lo), logpk, stack(rows_hi)])
    return np.concatenate([pad_lo, logk, pad_hi]), logpk


def _pad_log_knots(k, extrap_kmin=_default_extrap_kmin, extrap_kmax=_default_extrap_kmax):
    """log10(k) and the two continuation knots on each side (ref:62-67, 75-80): the range end and 90 % of the way to it."""
    logk = np.log10(np.asarray(k, dtype='f8'))
    lo = np.log10(min(extrap_kmin, k[0] * (1 - 1e-9)))
    hi = np.log10(max(extrap_kmax, k[-1] * (1 + 1e-9)))
    pad_hi = np.array([logk[-1] * 0.1 + hi * 0.9, hi])
    pad_lo = np.array([lo, logk[0] * 0.1 + lo * 0.9])
    return logk, pad_lo, pad_hi


def _transpose(a):
    """(n, B) <-> (B, n), contiguous."""
    if _buf.is_device_array(a):
        return a.T.contiguous()
    return np.ascontiguousarray(np.asarray(a).T)


class PowerSpectrumInterpolator1D(object):
    """
    1-D power-spectrum interpolator P(k) for one or many spectra sharing the k grid (``pk`` of shape (nk,) or
    (nk, ...)); same constructor and call signature as the reference's (ref:412-521).
    """

    def __init__(self, k, pk, interp_k='log', extrap_pk='log', extrap_kmin=_default_extrap_kmin, extrap_kmax=_default_extrap_kmax,
                 interp_order_k=3, device=None):
        self.k = np.asarray(k, dtype='f8').ravel()
        on_device = _buf.is_device_array(pk)
        self._pk = _buf.as_input(pk, dtype='f8').obj if on_device else np.asarray(pk, dtype='f8')
        ix = np.argsort(self.k)
        if not np.array_equal(ix, np.arange(self.k.size)):
            self.k = self.k[ix]
            self._pk = self._pk[ix] if not on_device else self._pk[_buf._torch().as_tensor(ix, device=self._pk.device)]
        self.interp_k, self.extrap_pk = str(interp_k), str(extrap_pk)
        self.interp_order_k = int(interp_order_k)
        self.extrap_kmin, self.extrap_kmax = self.k[0], self.k[-1]
        self._device = device
        self._rsigma8sq = 1.
        kk, pp = self.k, self._pk
        if self.extrap_pk == 'log':                                        # ref:343-351
            if self.interp_k != 'log':
                raise ValueError('log-log extrapolation requires log-x interpolation')
            self.extrap_kmin, self.extrap_kmax = extrap_kmin, extrap_kmax
            if self.interp_order_k != 3:
                raise NotImplementedError('cosmoprimo_b200 implements the cubic spline (interp_order_k=3) only')
            # logarithms, continuation knots (`_pad_log`), NaN screening and the fit in one pass over the table
            logk, pad_lo, pad_hi = _pad_log_knots(kk, extrap_kmin=extrap_kmin, extrap_kmax=extrap_kmax)
            self._interp = Interpolator1D.padlog(10**np.concatenate([pad_lo, logk, pad_hi]), pp, device=device)
            return
        self._interp = Interpolator1D(kk, pp, k=self.interp_order_k, interp_x=self.interp_k, interp_fun=self.extrap_pk,
                                      assume_sorted=True, device=device)

    @property
    def pk(self):
        return self._pk * self._rsigma8sq

    @property
    def kmin(self):
        return self.k[0]

    @property
    def kmax(self):
        return self.k[-1]

    def params(self):
        return dict(interp_k=self.interp_k, extrap_pk=self.extrap_pk, extrap_kmin=self.extrap_kmin, extrap_kmax=self.extrap_kmax,
                    interp_order_k=self.interp_order_k)

    def clone(self, **kwargs):
        """New interpolator with (possibly) other tables / settings (ref:366-373)."""
        state = dict(k=self.k, pk=self.pk, device=self._device, **self.params())
        state.update(kwargs)
        return self.__class__(**state)

    def __call__(self, k, **kwargs):
        """P(k); NaN outside [extrap_kmin, extrap_kmax]; shape ``k.shape + pk.shape[1:]`` (ref:495-521)."""
        return _scaled(self._interp(k, **kwargs), self._rsigma8sq)

    def sigma_r(self, r, nk=1024):
        r"""
        R.m.s. of perturbations in spheres of radius ``r``: FFTLog top-hat variance on ``nk`` log-spaced wavenumbers, then
        a natural cubic spline in (linear) s evaluated at ``r`` — ``integrate_sigma_r2(method='fftlog')``, ref:285-291.
        """
        rows = lambda k: (_scaled(self._interp.eval_rows(k), self._rsigma8sq), self._interp.shape)
        out = integrate_sigma_r2(r, self, kmin=self.extrap_kmin, kmax=self.extrap_kmax, nk=nk, device=self._device, pk_rows=rows, sqrt=True)
        if _buf.is_device_array(out):
            return out
        return out.astype(_bcast_dtype(r))

    def sigma8(self, **kwargs):
        return self.sigma_r(8., **kwargs)

    def rescale_sigma8(self, sigma8=1.):
        """Rescale the spectrum to the given sigma8 (ref:579-582)."""
        self._rsigma8sq = 1.
        self._rsigma8sq = sigma8**2 / self.sigma8()**2

    def to_xi(self, nk=1024, fftlog_kwargs=None, **kwargs):
        """Correlation function by FFTLog (ref:584-605)."""
        k = np.geomspace(self.extrap_kmin, self.extrap_kmax, nk)
        pk = self(k)
        lead = tuple(pk.shape[1:])
        fkw = dict(device=self._device)
        fkw.update(fftlog_kwargs or {})
        s, xi = PowerToCorrelation(k, complex=False, **fkw)(_transpose(pk.reshape(nk, -1)))
        params = dict(interp_s='log', interp_order_s=self.interp_order_k)
        params.update(kwargs)
        return CorrelationFunctionInterpolator1D(s, xi=_transpose(xi).reshape((nk,) + lead), device=self._device, **params)


class CorrelationFunctionInterpolator1D(object):
    """1-D correlation-function interpolator xi(s) (ref:1074-1215)."""

    def __init__(self, s, xi, interp_s='log', interp_order_s=3, device=None):
        self.s = np.asarray(s, dtype='f8').ravel()
        on_device = _buf.is_device_array(xi)
        self._xi = _buf.as_input(xi, dtype='f8').obj if on_device else np.asarray(xi, dtype='f8')
        ix = np.argsort(self.s)
        if not np.array_equal(ix, np.arange(self.s.size)):
            self.s = self.s[ix]
            self._xi = self._xi[ix] if not on_device else self._xi[_buf._torch().as_tensor(ix, device=self._xi.device)]
        self.interp_s = str(interp_s)
        self.interp_order_s = int(interp_order_s)
        self._device = device
        self._rsigma8sq = 1.
        self._interp = Interpolator1D(self.s, self._xi, k=self.interp_order_s, interp_x=self.interp_s, assume_sorted=True, device=device)

    @property
    def xi(self):
        return self._xi * self._rsigma8sq

    @property
    def smin(self):
        return self.s[0]

    @property
    def smax(self):
        return self.s[-1]

    extrap_smin, extrap_smax = smin, smax

    def __call__(self, s, **kwargs):
        return _scaled(self._interp(s, **kwargs), self._rsigma8sq)

    def to_pk(self, ns=1024, fftlog_kwargs=None, **kwargs):
        """Power spectrum by FFTLog (ref:1194-1215)."""
        s = np.geomspace(self.extrap_smin, self.extrap_smax, ns)
        xi = self(s)
        lead = tuple(xi.shape[1:])
        fkw = dict(device=self._device)
        fkw.update(fftlog_kwargs or {})
        k, pk = CorrelationToPower(s, complex=False, **fkw)(_transpose(xi.reshape(ns, -1)))
        params = dict(interp_k='log', interp_order_k=self.interp_order_s)
        params.update(kwargs)
        return PowerSpectrumInterpolator1D(k, pk=_transpose(pk).reshape((ns,) + lead), device=self._device, **params)

    def sigma_r(self, r, **kwargs):
        return self.to_pk().sigma_r(r, **kwargs)

    def sigma8(self, **kwargs):
        return self.sigma_r(8., **kwargs)


def integrate_sigma_r2(r, pk, kmin=1e-7, kmax=1e2, nk=None, device=None, pk_rows=None, sqrt=False):
    r"""
    Variance of perturbations in spheres of radius ``r``, :math:`\sigma_r^2 = \frac{1}{2\pi^2}\int dk\,k^2 P(k) W^2(kr)`, by
    the reference's default method (``integrate_sigma_r2(method='fftlog')``, ref:200, 285-291): FFTLog top-hat variance on
    ``nk`` (default 1024) log-spaced wavenumbers, then a natural cubic spline in (linear) s evaluated at ``r``.
    ``pk`` is a callable returning (nk,) or (nk, ...) for an array of wavenumbers; result ``r.shape + pk.shape[1:]``.
    ``pk_rows`` (optional, used by the interpolators of this module): callable returning the same values as (B, nk) rows
    and the trailing shape, so that nothing is transposed between the spline evaluation and FFTLog.  ``sqrt``: return sigma_r.
    """
    if nk is None: nk = 1024
    k = np.geomspace(kmin, kmax, nk)
    if pk_rows is not None:
        rows, lead = pk_rows(k)
        dtype = _bcast_dtype(r) if not lead else np.dtype('f8')
    else:
        p = pk(k)
        lead = tuple(p.shape[1:])
        dtype = _bcast_dtype(r, p if p.ndim > 1 else None)
        rows = _transpose(p.reshape(nk, -1))
    rr = np.asarray(r, dtype='f8')
    s, var = TophatVariance(k, device=device)(rows)                                        # (B, nk)
    if _buf.is_device_array(var):
        # device rows: the reference's factor 2 pi^2 (ref:289) and its inverse (ref:292) cancel; no pass over the result, and the
        # square root of sigma_r (``sqrt``) is taken by the kernel that writes it
        sigma2 = spline_eval_rows(s, var, rr.ravel(), device=device, sqrt=sqrt).reshape(rr.shape + lead)
    else:
        tmp = (2. * np.pi**2) * spline_eval_rows(s, var, rr.ravel(), device=device)            # ref:289, rows layout
        sigma2 = 1. / (2. * np.pi**2) * tmp.reshape(rr.shape + lead)
        if sqrt:
            sigma2 = sigma2**0.5
    if _buf.is_device_array(sigma2):
        return sigma2.to(_buf._torch().float32) if dtype == np.float32 else sigma2
    return sigma2.astype(dtype)


def _scaled(a, factor):
    """a * factor, without a pass over the array when the factor is exactly 1 (no sigma8 rescaling requested)."""
    return a if (np.ndim(factor) == 0 and factor == 1.) else a * factor


def _times(a, b):
    """a * b where ``a`` may be a torch CUDA tensor and ``b`` a numpy array (broadcast over the last axis)."""
    if _buf.is_device_array(a) and isinstance(b, np.ndarray):
        b = _buf._torch().as_tensor(b, device=a.device)
    return a * b


class PowerSpectrumInterpolator2D(object):
    """
    2-D power-spectrum interpolator P(k, z) (ref:608-987): bicubic spline of log10 P in (log10 k, z) with the log-log
    extrapolation knots of ``_pad_log``, optional ``growth_factor_sq`` callable.  ``pk`` has shape (nk, nz); with a single
    redshift column ``growth_factor_sq`` carries the z dependence (ref:668-672).
    """

    def __init__(self, k, z, pk, interp_k='log', extrap_pk='log', extrap_kmin=_default_extrap_kmin, extrap_kmax=_default_extrap_kmax,
                 interp_order_k=3, interp_order_z=3, growth_factor_sq=None, device=None):
        self._rsigma8sq = 1.
        self.growth_factor_sq = growth_factor_sq
        self._device = device
        self.k = np.asarray(k, dtype='f8').ravel()
        self.z = np.asarray(z, dtype='f8').ravel()
        on_device = _buf.is_device_array(pk)
        self._pk = (_buf.as_input(pk, dtype='f8').obj if on_device else np.asarray(pk, dtype='f8')).reshape(self.k.size, -1)
        ik, iz = np.argsort(self.k), np.argsort(self.z)
        if not (np.array_equal(ik, np.arange(ik.size)) and np.array_equal(iz, np.arange(iz.size))):
            self.k, self.z = self.k[ik], self.z[iz]
            if on_device:
                torch = _buf._torch()
                self._pk = self._pk[torch.as_tensor(ik, device=self._pk.device)][:, torch.as_tensor(iz, device=self._pk.device)]
            else:
                self._pk = self._pk[np.ix_(ik, iz)]
        self.interp_k, self.extrap_pk = str(interp_k), str(extrap_pk)
        self.interp_order_k, self.interp_order_z = int(interp_order_k), int(interp_order_z)
        kk, pp = self.k, self._pk
        self.extrap_kmin, self.extrap_kmax = self.k[0], self.k[-1]
        if self.extrap_pk == 'log':                                        # ref:343-351
            if self.interp_k != 'log':
                raise ValueError('log-log extrapolation requires log-x interpolation')
            self.extrap_kmin, self.extrap_kmax = extrap_kmin, extrap_kmax
            kk, pp = _pad_log(kk, pp, extrap_kmin=extrap_kmin, extrap_kmax=extrap_kmax)
            kk, pp = 10**kk, 10**pp
        self._is2d = self._pk.shape[1] > 1
        if self._is2d:
            self._interp = Interpolator2D(kk, self.z, pp, kx=self.interp_order_k, ky=self.interp_order_z, interp_x=self.interp_k,
                                          interp_fun=self.extrap_pk, assume_sorted=True, device=device)
        else:
            if self.growth_factor_sq is None:
                raise ValueError('provide either 2D pk array or growth_factor_sq')
            self._interp = Interpolator1D(kk, pp[:, 0], k=self.interp_order_k, interp_x=self.interp_k, interp_fun=self.extrap_pk,
                                          assume_sorted=True, device=device)

    @property
    def pk(self):
        return self._pk * self._rsigma8sq

    kmin = property(lambda self: self.k[0])
    kmax = property(lambda self: self.k[-1])
    zmin = property(lambda self: self.z[0])
    zmax = property(lambda self: self.z[-1])

    def params(self):
        return dict(interp_k=self.interp_k, extrap_pk=self.extrap_pk, extrap_kmin=self.extrap_kmin, extrap_kmax=self.extrap_kmax,
                    interp_order_k=self.interp_order_k, interp_order_z=self.interp_order_z, growth_factor_sq=self.growth_factor_sq)

    def clone(self, **kwargs):
        state = dict(k=self.k, z=self.z, pk=self.pk, device=self._device, **self.params())
        state.update(kwargs)
        return self.__class__(**state)

    def __call__(self, k, z, grid=True, ignore_growth=False, bounds_error=False, rows=False):
        """P(k, z): shape ``k.shape + z.shape`` if ``grid`` else ``k.shape``; NaN outside the (extrapolated) ranges (ref:720-800).
        ``rows=True`` (grid, 1-D arguments): the float64 result transposed, (nz, nk), one row per redshift (FFTLog's layout)."""
        dtype = _bcast_dtype(k, z)
        k, z = (np.asarray(xx, dtype=dtype) for xx in (k, z))
        shape = k.shape + z.shape if grid else k.shape
        k, z = k.ravel(), z.ravel()
        rows = bool(rows and grid)
        mask_k = (k >= self.extrap_kmin) & (k <= self.extrap_kmax)
        mask_z = (z >= self.zmin) & (z <= self.zmax)
        if bounds_error and not (mask_k.all() and (mask_z.all() or not self._is2d)):
            raise ValueError('input outside of extrapolation range')
        if not self._is2d: mask_z = np.ones_like(mask_z)                    # ignore input z (ref:784)
        mask = mask_k[:, None] & mask_z if grid else mask_k & mask_z
        if self._is2d:
            # queries outside [zmin, zmax] are masked below; the bicubic table clamps them to its edge meanwhile
            tmp = self._interp(k, z, grid=grid, rows=rows)
        else:
            tmp = self._interp(k)
            if grid:
                if rows:
                    tmp = tmp[None, :].expand(z.size, -1) if _buf.is_device_array(tmp) else np.repeat(tmp[None, :], z.size, axis=0)
                else:
                    tmp = tmp[:, None].expand(-1, z.size) if _buf.is_device_array(tmp) else np.repeat(tmp[:, None], z.size, axis=-1)
        if rows:
            mask, shape, dtype = mask.T, (z.size, k.size), np.dtype('f8')
        if self.growth_factor_sq is not None and not ignore_growth:
            growth = np.asarray(self.growth_factor_sq(z)).astype(dtype)
            tmp = _times(tmp, growth[:, None] if rows else growth)
        if _buf.is_device_array(tmp):
            torch = _buf._torch()
            tmp = torch.where(torch.as_tensor(mask, device=tmp.device), tmp, torch.full_like(tmp, float('nan')))
            return _scaled((tmp.to(torch.float32 if dtype == np.float32 else torch.float64)).reshape(shape), self._rsigma8sq)
        return _scaled(np.where(mask, tmp, np.nan).astype(dtype).reshape(shape), self._rsigma8sq)

    def sigma_rz(self, r, z, nk=None):
        """R.m.s. of perturbations in spheres of radius ``r`` at redshifts ``z``: (r.size, z.size) (ref:846-876)."""
        # redshifts outside the table give NaN columns in the reference (ref:781-795).  The FFTLog kernels transform rows in
        # pairs (two real rows = one complex FFT), so a NaN row would also spoil its partner: evaluate those rows at the
        # clipped redshift and blank the result afterwards instead.
        zz = np.asarray(z, dtype='f8')
        bad = ~((zz >= self.zmin) & (zz <= self.zmax)) if self._is2d else np.zeros(zz.shape, dtype='?')
        zc = np.clip(zz, self.zmin, self.zmax) if self._is2d else zz
        rows = lambda k: (self(k, zc.ravel(), rows=True), zz.shape)
        toret = integrate_sigma_r2(r, lambda k: self(k, zc), kmin=self.extrap_kmin, kmax=self.extrap_kmax, nk=nk, device=self._device, pk_rows=rows, sqrt=True)
        if bad.any():
            toret[..., _buf._torch().as_tensor(bad, device=toret.device) if _buf.is_device_array(toret) else bad] = float('nan')
        dtype = _bcast_dtype(r, z)
        if _buf.is_device_array(toret):
            return toret.to(_buf._torch().float32) if dtype == np.float32 else toret
        return toret.astype(dtype)

    def sigma8_z(self, z=0, **kwargs):
        return self.sigma_rz(8., z=z, **kwargs)

    def rescale_sigma8(self, sigma8=1.):
        """Rescale to the given sigma8 at z = 0 (ref:881-884)."""
        self._rsigma8sq = 1.
        self._rsigma8sq = sigma8**2 / float(np.asarray(_to_host(self.sigma8_z(z=0))))**2

    def growth_rate_rz(self, r, z, dz=1e-3, **kwargs):
        """f(r, z) = d ln sigma_r / d ln a by finite differences of ``sigma_rz`` (ref:886-936): the five shifted redshift
        grids go through ONE FFTLog batch."""
        if self.interp_order_z == 0 and self.growth_factor_sq is None:
            import warnings
            warnings.warn('No redshift evolution provided, growth rate is 0')
            return 0.
        hdz = dz / 2.
        dtype = _bcast_dtype(r, z)
        r, z = (np.asarray(xx, dtype=dtype) for xx in (r, z))
        shape = r.shape + z.shape
        if not all(shape):
            return np.zeros(shape, dtype=dtype)
        z = z.ravel()
        zall = np.concatenate([z - dz, z - hdz, z, z + hdz, z + dz])
        sig = _to_host(self.sigma_rz(r.ravel(), zall, **kwargs)).astype('f8').reshape(-1, 5, z.size)
        feval = np.log(sig)
        toret = np.where(z < self.zmin + hdz, -feval[:, 4] + 4 * feval[:, 3] - 3 * feval[:, 2], feval[:, 3] - feval[:, 1])
        toret = np.where(z > self.zmax - hdz, -(-feval[:, 0] + 4 * feval[:, 1] - 3 * feval[:, 2]), toret)
        dsigdlna = -(toret / dz) * (1 + z)
        return dsigdlna.astype(dtype).reshape(shape)

    def to_1d(self, z, **kwargs):
        """:class:`PowerSpectrumInterpolator1D` at redshift(s) ``z`` (ref:938-963)."""
        params = dict(extrap_pk=self.extrap_pk, extrap_kmin=self.extrap_kmin, extrap_kmax=self.extrap_kmax, interp_order_k=self.interp_order_k,
                      device=self._device)
        params.update(kwargs)
        saved = self.extrap_kmin, self.extrap_kmax
        self.extrap_kmin, self.extrap_kmax = -np.inf, np.inf               # in case self.k > self.extrap_kmax (ref:959)
        try:
            pk = self(self.k, z=z)
        finally:
            self.extrap_kmin, self.extrap_kmax = saved
        return PowerSpectrumInterpolator1D(self.k, pk, **params)

    def to_xi(self, nk=1024, fftlog_kwargs=None, **kwargs):
        """Correlation function by FFTLog (ref:965-987)."""
        k = np.geomspace(self.extrap_kmin, self.extrap_kmax, nk)
        fkw = dict(device=self._device)
        fkw.update(fftlog_kwargs or {})
        s, xi = PowerToCorrelation(k, complex=False, **fkw)(_transpose(self(k, z=self.z, ignore_growth=True)))
        params = dict(interp_s='log', interp_order_s=self.interp_order_k, interp_order_z=self.interp_order_z,
                      growth_factor_sq=self.growth_factor_sq, device=self._device)
        params.update(kwargs)
        return CorrelationFunctionInterpolator2D(s, z=self.z, xi=_transpose(xi), **params)


def _to_host(a):
    return a.cpu().numpy() if _buf.is_device_array(a) else np.asarray(a)


class CorrelationFunctionInterpolator2D(object):
    """2-D correlation-function interpolator xi(s, z) (ref:1219-1498): bicubic spline of xi in (log10 s, z)."""

    def __init__(self, s, z, xi=None, interp_s='log', interp_order_s=3, interp_order_z=3, growth_factor_sq=None, device=None):
        self._rsigma8sq = 1.
        self.growth_factor_sq = growth_factor_sq
        self._device = device
        self.s = np.asarray(s, dtype='f8').ravel()
        self.z = np.asarray(z, dtype='f8').ravel()
        on_device = _buf.is_device_array(xi)
        self._xi = (_buf.as_input(xi, dtype='f8').obj if on_device else np.asarray(xi, dtype='f8')).reshape(self.s.size, -1)
        i_s, iz = np.argsort(self.s), np.argsort(self.z)
        if not (np.array_equal(i_s, np.arange(i_s.size)) and np.array_equal(iz, np.arange(iz.size))):
            self.s, self.z = self.s[i_s], self.z[iz]
            if on_device:
                torch = _buf._torch()
                self._xi = self._xi[torch.as_tensor(i_s, device=self._xi.device)][:, torch.as_tensor(iz, device=self._xi.device)]
            else:
                self._xi = self._xi[np.ix_(i_s, iz)]
        self.interp_s = str(interp_s)
        self.interp_order_s, self.interp_order_z = int(interp_order_s), int(interp_order_z)
        self._is2d = self._xi.shape[1] > 1
        if self._is2d:
            self._interp = Interpolator2D(self.s, self.z, self._xi, kx=self.interp_order_s, ky=self.interp_order_z, interp_x=self.interp_s,
                                          assume_sorted=True, device=device)
        else:
            if self.growth_factor_sq is None:
                raise ValueError('provide either 2D xi array or growth_factor_sq')
            self._interp = Interpolator1D(self.s, self._xi[:, 0], k=self.interp_order_s, interp_x=self.interp_s, assume_sorted=True, device=device)

    @property
    def xi(self):
        return self._xi * self._rsigma8sq

    smin = property(lambda self: self.s[0])
    smax = property(lambda self: self.s[-1])
    extrap_smin, extrap_smax = smin, smax
    zmin = property(lambda self: self.z[0])
    zmax = property(lambda self: self.z[-1])

    def __call__(self, s, z, grid=True, ignore_growth=False, bounds_error=False):
        """xi(s, z) (ref:1337-1414)."""
        dtype = _bcast_dtype(s, z)
        s, z = (np.asarray(xx, dtype=dtype) for xx in (s, z))
        shape = s.shape + z.shape if grid else s.shape
        s, z = s.ravel(), z.ravel()
        mask_s = (s >= self.smin) & (s <= self.smax)
        mask_z = (z >= self.zmin) & (z <= self.zmax)
        if bounds_error and not (mask_s.all() and (mask_z.all() or not self._is2d)):
            raise ValueError('input outside of extrapolation range')
        if not self._is2d: mask_z = np.ones_like(mask_z)
        mask = mask_s[:, None] & mask_z if grid else mask_s & mask_z
        if self._is2d:
            tmp = self._interp(np.clip(s, self.smin, self.smax), np.clip(z, self.zmin, self.zmax), grid=grid)
        else:
            tmp = self._interp(np.clip(s, self.smin, self.smax))
            if grid:
                tmp = tmp[:, None].expand(-1, z.size) if _buf.is_device_array(tmp) else np.repeat(tmp[:, None], z.size, axis=-1)
        if self.growth_factor_sq is not None and not ignore_growth:
            tmp = _times(tmp, np.asarray(self.growth_factor_sq(z)).astype(dtype))
        if _buf.is_device_array(tmp):
            torch = _buf._torch()
            tmp = torch.where(torch.as_tensor(mask, device=tmp.device), tmp, torch.full_like(tmp, float('nan')))
            return _scaled((tmp.to(torch.float32 if dtype == np.float32 else torch.float64)).reshape(shape), self._rsigma8sq)
        return _scaled(np.where(mask, tmp, np.nan).astype(dtype).reshape(shape), self._rsigma8sq)

    def to_pk(self, ns=1024, fftlog_kwargs=None, **kwargs):
        """Power spectrum by FFTLog (ref:1476-1498)."""
        s = np.geomspace(self.extrap_smin, self.extrap_smax, ns)
        fkw = dict(device=self._device)
        fkw.update(fftlog_kwargs or {})
        k, pk = CorrelationToPower(s, complex=False, **fkw)(_transpose(self(s, self.z, ignore_growth=True)))
        params = dict(interp_k='log', extrap_pk='log', interp_order_k=self.interp_order_s, interp_order_z=self.interp_order_z,
                      growth_factor_sq=self.growth_factor_sq, device=self._device)
        params.update(kwargs)
        return PowerSpectrumInterpolator2D(k, z=self.z, pk=_transpose(pk), **params)

    def to_1d(self, z, **kwargs):
        """:class:`CorrelationFunctionInterpolator1D` at redshift(s) ``z`` (ref:1453-1474)."""
        params = dict(interp_order_s=self.interp_order_s, device=self._device)
        params.update(kwargs)
        return CorrelationFunctionInterpolator1D(self.s, self(self.s, z=z), **params)

    def sigma_rz(self, r, z, **kwargs):
        return self.to_pk().sigma_rz(r, z=z, **kwargs)

    def sigma8_z(self, z, **kwargs):
        return self.sigma_rz(8., z=z, **kwargs)

    def growth_rate_rz(self, r, z, **kwargs):
        return self.to_pk().growth_rate_rz(r, z=z, **kwargs)
