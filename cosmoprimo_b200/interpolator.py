"""
The callers either side of the FFTLog path, GPU-backed: 1-D power-spectrum / correlation-function interpolators with
the reference's call signatures (``cosmoprimo/interpolator.py``, cited as ``ref:LINE``).  Only what sits on the hot
path is here: log-log extrapolation padding (``_pad_log``, ref:42-87), spline evaluation (ref:495-521),
``sigma_r`` with ``method='fftlog'`` (ref:200-292, branch 285-289) and ``to_xi`` / ``to_pk`` (ref:584-605,
1194-1215).  Cloning/pytree/2-D (k, z) plumbing stays in the reference.

numpy tables in -> numpy results; CUDA tables in -> torch results (the spline lives on the device either way).
"""

import numpy as np

from . import _buffers as _buf
from .interp import Interpolator1D, _bcast_dtype, spline_eval_rows
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
    logk = np.log10(np.asarray(k, dtype='f8'))
    logpk = xp.log10(pk)
    lo = np.log10(min(extrap_kmin, k[0] * (1 - 1e-9)))
    hi = np.log10(max(extrap_kmax, k[-1] * (1 + 1e-9)))
    pad_hi = np.array([logk[-1] * 0.1 + hi * 0.9, hi])
    pad_lo = np.array([lo, logk[0] * 0.1 + lo * 0.9])
    slope_hi = (logpk[-1] - logpk[-2]) / (logk[-1] - logk[-2])
    slope_lo = (logpk[1] - logpk[0]) / (logk[1] - logk[0])
    rows_hi = [logpk[-1] + slope_hi * (x - logk[-1]) for x in pad_hi]
    rows_lo = [logpk[0] + slope_lo * (x - logk[0]) for x in pad_lo]
    stack = xp.stack if xp is not np else np.stack
    cat = xp.cat if xp is not np else np.concatenate
    logpk = cat([stack(rows_lo), logpk, stack(rows_hi)])
    return np.concatenate([pad_lo, logk, pad_hi]), logpk


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
            kk, pp = _pad_log(kk, pp, extrap_kmin=extrap_kmin, extrap_kmax=extrap_kmax)
            kk, pp = 10**kk, 10**pp
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
        return self._interp(k, **kwargs) * self._rsigma8sq

    def sigma_r(self, r, nk=1024):
        r"""
        R.m.s. of perturbations in spheres of radius ``r``: FFTLog top-hat variance on ``nk`` log-spaced wavenumbers, then
        a natural cubic spline in (linear) s evaluated at ``r`` — ``integrate_sigma_r2(method='fftlog')``, ref:285-291.
        """
        k = np.geomspace(self.extrap_kmin, self.extrap_kmax, nk)
        pk = self(k)
        lead = tuple(pk.shape[1:])
        s, var = TophatVariance(k, device=self._device)(_transpose(pk.reshape(nk, -1)))      # (B, nk)
        # natural spline of var in (linear) s evaluated at r, ref:289 -- on the (B, nk) rows as FFTLog wrote them
        dtype = _bcast_dtype(r, pk if pk.ndim > 1 else None)
        rr = np.asarray(r, dtype='f8')
        tmp = (2. * np.pi**2) * spline_eval_rows(s, var, rr.ravel(), device=self._device)
        sigma2 = 1. / (2. * np.pi**2) * tmp.reshape(rr.shape + lead)
        out = sigma2**0.5
        if _buf.is_device_array(out):
            return out.to(_buf._torch().float32) if dtype == np.float32 else out
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
        return self._interp(s, **kwargs) * self._rsigma8sq

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
