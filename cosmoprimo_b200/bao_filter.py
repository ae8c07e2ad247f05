"""
Wallish2018 BAO no-wiggle filter on the GPU: plug-in for the reference's filter registry
(``cosmoprimo/bao_filter.py``, cited as ``ref:LINE``: metaclass registry 22-31, base class 34-169, Wallish2018 345-431,
factory 912-921).

The class keeps the reference's contract — ``name``, ``set_k`` / ``set_pk`` / ``_prepare`` / ``_compute``,
attributes ``k``, ``pk``, ``pknow``, ``wiggles``, ``smooth_pk_interpolator()`` — and accepts the reference's own
interpolator objects (anything callable with ``extrap_kmin`` / ``extrap_kmax``) as well as
:class:`cosmoprimo_b200.interpolator.PowerSpectrumInterpolator1D`.  ``_compute`` makes the same two evaluations of the
input interpolator as the reference (ref:364-369 and 92-102) and hands them to ``cpf_wallish2018``: DST-II, the
even/odd spline second derivatives, the argmax boxes, cut + re-spline, DST-III, the splice and the final clamped spline
all run on the device (``csrc/cpf_wallish.cu``).
"""

import numpy as np

from . import _lib
from . import _buffers as _buf

_registry = {}


def register(cls):
    """Register a filter class under ``cls.name`` (the reference does it with a metaclass, ref:22-31)."""
    _registry[cls.name] = cls
    return cls


class BasePowerSpectrumBAOFilter(object):
    """Base BAO filter for the power spectrum (ref:34-169)."""
    name = 'base'

    def __init__(self, pk_interpolator, cosmo=None, cosmo_fid=None, device=None, **kwargs):
        self._cosmo_fid = cosmo_fid
        self._cosmo = cosmo
        self._device = device
        self.pk_interpolator = pk_interpolator
        self.set_k(**kwargs)
        self.set_pk(pk_interpolator, cosmo=cosmo)
        self._prepare()
        self._compute()
        self.pk, self.pknow = (x.reshape(self.shape) for x in (self.pk, self.pknow))

    def _prepare(self):
        """Anything that can be done once."""

    def set_k(self, nk=1024):
        """Wavenumbers of the output, log-spaced over the extrapolation range of the interpolator (ref:81-90)."""
        self.k = np.geomspace(self.pk_interpolator.extrap_kmin, self.pk_interpolator.extrap_kmax, nk)

    def _evaluate(self, k):
        """pk_interpolator(k) with the (k, z) form of 2-D interpolators (ref:96-99, 365-368), as (nk, ncols)."""
        interp = self.pk_interpolator
        if hasattr(interp, 'z') and getattr(interp, 'z', None) is not None and np.ndim(interp.z) > 0 and hasattr(interp, 'growth_factor_sq'):
            pk = interp(k, interp.z, ignore_growth=True)
        else:
            pk = interp(k)
        return pk

    def set_pk(self, pk_interpolator, cosmo=None):
        """Evaluate the input spectrum on :attr:`k` (ref:92-102)."""
        if cosmo is not None: self._cosmo = cosmo
        self.pk_interpolator = pk_interpolator
        self.pk = self._evaluate(self.k)
        self.shape = tuple(self.pk.shape)
        self.pk = self.pk.reshape(self.pk.shape[0], -1)

    def __call__(self, pk_interpolator, cosmo=None):
        """Re-run with a new interpolator (ref:104-108)."""
        self.set_pk(pk_interpolator, cosmo=cosmo)
        self._compute()
        self.pk, self.pknow = (x.reshape(self.shape) for x in (self.pk, self.pknow))
        return self

    @property
    def wiggles(self):
        """Extracted wiggles."""
        return self.pk / self.pknow

    def smooth_pk_interpolator(self, **kwargs):
        """Smooth (no-wiggle) power spectrum interpolator (ref:115-129)."""
        return self.pk_interpolator.clone(k=self.k, pk=self.pknow, **kwargs)

    def smooth_xi_interpolator(self, **kwargs):
        """Smooth (no-peak) correlation function through FFTLog (ref:131-145)."""
        return self.smooth_pk_interpolator().to_xi(**kwargs)


@register
class Wallish2018PowerSpectrumBAOFilter(BasePowerSpectrumBAOFilter):
    """
    Filter BAO wiggles by sine-transforming log(k P) to real space, cutting the bump in the even and odd coefficients and
    re-interpolating with splines (ref:345-431; arXiv:1810.02800 appendix D, arXiv:1003.3999), on a CUDA device.
    """
    name = 'wallish2018_cuda'

    def _compute(self):
        lib = _lib.load()
        _lib.require_device()
        klin = np.linspace(self.pk_interpolator.extrap_kmin, 2., 4096)                       # ref:364
        pklin = self._evaluate(klin)                                                          # ref:365-369
        pklin = pklin.reshape(pklin.shape[0], -1)
        lin, out_in = _buf.as_input(pklin, dtype='f8'), _buf.as_input(self.pk, dtype='f8')
        if lin.on_device != out_in.on_device:
            raise ValueError('pk_interpolator returned host and device arrays for the two grids')
        ncols = int(lin.shape[1])
        if ncols != int(out_in.shape[1]):
            raise ValueError('pk_interpolator returned {} and {} spectra on the two grids'.format(ncols, out_in.shape[1]))
        device = lin.device if lin.on_device else (self._device if self._device is not None else _buf.default_device())
        if lin.on_device:
            torch = _buf._torch()
            dev = torch.device('cuda', device)
            kl, ko = _buf.as_input(torch.as_tensor(klin, device=dev)), _buf.as_input(torch.as_tensor(self.k, device=dev))
            boxes = torch.empty((ncols, 4), dtype=torch.int32, device=dev)
            boxes_ptr = boxes.data_ptr()
            stream = _buf.current_stream(device)
        else:
            kl, ko = _buf.as_input(klin), _buf.as_input(self.k)
            boxes = np.empty((ncols, 4), dtype='i4')
            boxes_ptr = boxes.ctypes.data
            stream = None
        res = _buf.empty_like_kind(out_in, (self.k.size, ncols), dtype='f8')
        rc = lib.cpf_wallish2018(kl.ptr, lin.ptr, 4096, ko.ptr, out_in.ptr, self.k.size, ncols, res.ptr, boxes_ptr,
                                 int(lin.on_device), device, stream)
        _lib.check(rc)
        self.pknow = res.obj
        self.pk = out_in.obj
        # ibox_even, ibox_odd of every column (ref:394-395), kept for inspection like the reference's _dd_* attributes
        self._boxes = boxes


def PowerSpectrumBAOFilter(pk_interpolator, engine='wallish2018_cuda', **kwargs):
    """Factory (ref:912-921): ``engine`` is one of the registered names; 'wallish2018' is accepted as an alias."""
    name = engine.lower()
    if name == 'wallish2018':
        name = 'wallish2018_cuda'
    try:
        cls = _registry[name]
    except KeyError:
        raise ValueError('Power spectrum BAO filter {} is unknown; cosmoprimo_b200 provides {}'.format(engine, sorted(_registry)))
    return cls(pk_interpolator, **kwargs)


def register_in_reference():
    """
    Make ``cosmoprimo.bao_filter.PowerSpectrumBAOFilter(interp, engine='wallish2018_cuda')`` work with the unmodified
    reference by adding this class to its registry (``RegisteredPowerSpectrumBAOFilter._registry``, ref:22-31).
    """
    from cosmoprimo import bao_filter as ref
    ref.RegisteredPowerSpectrumBAOFilter._registry[Wallish2018PowerSpectrumBAOFilter.name] = Wallish2018PowerSpectrumBAOFilter
    return ref
