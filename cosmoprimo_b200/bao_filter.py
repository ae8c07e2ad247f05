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

    def _evaluate_rows(self, k):
        """The same values with one ROW per spectrum, (ncols, nk), written directly by the spline kernel (``cpf_spline_eval_t``) -- the layout
        ``cpf_wallish2018_rows`` fetches with bulk copies -- or None when the interpolator is not one of this package's."""
        from .interpolator import PowerSpectrumInterpolator1D, PowerSpectrumInterpolator2D, _scaled
        interp = self.pk_interpolator
        if type(interp) is PowerSpectrumInterpolator1D:
            return _scaled(interp._interp.eval_rows(k), interp._rsigma8sq)
        if type(interp) is PowerSpectrumInterpolator2D and interp._is2d and np.ndim(interp.z) > 0:
            return interp(k, interp.z, ignore_growth=True, rows=True)
        return None

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
        pklin = self._evaluate_rows(klin)                                                     # ref:365-369, one row per spectrum where the interpolator can
        rows = pklin is not None
        if rows:
            pklin = pklin.reshape(-1, klin.size)
        else:
            pklin = self._evaluate(klin)
            pklin = pklin.reshape(pklin.shape[0], -1)
        lin, out_in = _buf.as_input(pklin, dtype='f8'), _buf.as_input(self.pk, dtype='f8')
        if lin.on_device != out_in.on_device:
            raise ValueError('pk_interpolator returned host and device arrays for the two grids')
        ncols = int(lin.shape[0 if rows else 1])
        if ncols != int(out_in.shape[1]):
            raise ValueError('pk_interpolator returned {} and {} spectra on the two grids'.format(ncols, out_in.shape[1]))
        device = lin.device if lin.on_device else (self._device if self._device is not None else _buf.default_device())
        if lin.on_device:
            torch = _buf._torch()
            dev = torch.device('cuda', device)
            if rows:          # the rows entry takes the two grids as host arrays (cached plan: no copy, no synchronisation)
                kl, ko = _buf.as_input(klin), _buf.as_input(self.k)
            else:
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
        entry = lib.cpf_wallish2018_rows if rows else lib.cpf_wallish2018
        rc = entry(kl.ptr, lin.ptr, 4096, ko.ptr, out_in.ptr, self.k.size, ncols, res.ptr, boxes_ptr, int(lin.on_device), device, stream)
        _lib.check(rc)
        self.pknow = res.obj
        self.pk = out_in.obj
        # ibox_even, ibox_odd of every column (ref:394-395), kept for inspection like the reference's _dd_* attributes
        self._boxes = boxes


# ---- the filters that are least squares + splines -----------------------------------------------------------------
# Cosmology-dependent inputs (one cosmology per filter, host side): `cosmo.rs_drag` and the Eisenstein & Hu no-wiggle spectrum.  Accepted
# objects: cosmoprimo_b200.eisenstein_hu.EHCosmology (or anything with `rs_drag`, `pk_nowiggle(k)`, `pk_lin(k)`), or the reference's own
# Cosmology (the same calls the reference makes, ref:320, 528-532, 576).

def _xp(arr):
    """numpy for host arrays, torch for CUDA tensors."""
    return np if isinstance(arr, np.ndarray) else _buf._torch()


def _like(vec, arr):
    """Host vector ``vec`` as the same kind of array as ``arr`` (numpy, or a CUDA tensor on arr's device)."""
    if isinstance(arr, np.ndarray):
        return np.asarray(vec, dtype='f8')
    return _buf._torch().as_tensor(np.asarray(vec, dtype='f8'), device=arr.device)


def _on_device(arr, device=None):
    """(CUDA tensor, was_numpy): host arrays are moved to the device -- these filters compute on the GPU only, like the rest of the package
    (no CPU fallback: without a device this raises)."""
    if not isinstance(arr, np.ndarray):
        return arr, False
    _lib.require_device()
    torch = _buf._torch()
    return torch.as_tensor(np.ascontiguousarray(arr, dtype='f8'), device=torch.device('cuda', device if device is not None else _buf.default_device())), True


def _pk_nowiggle(cosmo, k):
    if hasattr(cosmo, 'pk_nowiggle'):
        return np.asarray(cosmo.pk_nowiggle(k), dtype='f8')
    from cosmoprimo.cosmology import Fourier          # a reference Cosmology object: the reference's own call (ref:320)
    return np.asarray(Fourier(cosmo, engine='eisenstein_hu_nowiggle', set_engine=False).pk_interpolator()(k, z=0.), dtype='f8')


def _pk_lin(cosmo, k):
    if hasattr(cosmo, 'pk_lin'):
        return np.asarray(cosmo.pk_lin(k), dtype='f8')
    from cosmoprimo.cosmology import Fourier          # ref:528
    return np.asarray(Fourier(cosmo).pk_interpolator()(k, z=0.), dtype='f8')


class _CosmoMixin(object):
    """``cosmo`` / ``cosmo_fid`` / ``rs_drag_ratio`` of the reference's base classes (ref:147-169)."""
    RS_DRAG_FID = 100.91463132327911          # ref:166: rs_drag of the reference's default Cosmology()

    def rs_drag_ratio(self):
        if self._cosmo is None:
            return 1.
        rs_drag_fid = self.RS_DRAG_FID if self._cosmo_fid is None else self._cosmo_fid.rs_drag
        return float(self._cosmo.rs_drag / rs_drag_fid)

    @property
    def cosmo(self):
        if self._cosmo is None:
            raise ValueError('cosmo must be provided (an EHCosmology or a reference Cosmology)')
        return self._cosmo

    @property
    def cosmo_fid(self):
        if self._cosmo_fid is None:
            raise ValueError('cosmo_fid must be provided, with an engine')
        return self._cosmo_fid


@register
class EHNoWigglePolyPowerSpectrumBAOFilter(_CosmoMixin, BasePowerSpectrumBAOFilter):
    """Remove BAO wiggles with the Eisenstein & Hu no-wiggle formula corrected by a constrained polynomial k^-2 .. k^3 (ref:289-342); the fit
    of all spectra is one matrix product with the pre-solved least-squares operator."""
    name = 'ehpoly_cuda'

    def __init__(self, pk_interpolator, krange=(1e-3, 1.), rescale_krange=True, cosmo=None, **kwargs):
        self.krange = krange
        self.rescale_krange = rescale_krange
        super(EHNoWigglePolyPowerSpectrumBAOFilter, self).__init__(pk_interpolator, cosmo=cosmo, **kwargs)

    def _compute(self):
        from .utils import LeastSquareSolver
        krange = np.asarray(self.krange, dtype='f8')
        if self.rescale_krange:
            krange = krange / self.rs_drag_ratio()                                                     # ref:316-317
        mask = (self.k >= krange[0]) & (self.k <= krange[1])
        k = self.k[mask]
        pk, was_numpy = _on_device(self.pk, self._device)
        xp = _xp(pk)
        lo, hi = int(np.flatnonzero(mask)[0]), int(np.flatnonzero(mask)[-1]) + 1                        # the mask is one contiguous range
        ratio = (pk[lo:hi] / _like(_pk_nowiggle(self.cosmo, k), pk)[:, None]).T.contiguous()             # ref:320, (ncols, nk_mask)
        gradient = np.array([k**(i - 2) for i in range(6)])                                             # ref:322
        constraint_gradient = np.column_stack([gradient[..., 0], gradient[..., 1] - gradient[..., 0], gradient[..., -1], gradient[..., -2] - gradient[..., -1]])
        solver = LeastSquareSolver(gradient, precision=k**2, constraint_gradient=constraint_gradient, compute_inverse=False)
        solver(ratio, constraint=xp.stack([ratio[..., 0], ratio[..., 1] - ratio[..., 0], ratio[..., -1], ratio[..., -2] - ratio[..., -1]], dim=-1))   # ref:325
        wiggles = xp.ones_like(pk)
        wiggles[lo:hi] = (ratio / solver.model()).T                                                     # ref:327-328
        pknow = pk / wiggles
        self.pknow = pknow.cpu().numpy() if was_numpy else pknow


@register
class PeakAveragePowerSpectrumBAOFilter(_CosmoMixin, BasePowerSpectrumBAOFilter):
    r"""Average of the splines through the maxima and through the minima of the wiggles, at the fiducial peak positions rescaled by
    :math:`r_{\mathrm{drag}} / r_{\mathrm{drag}}^{\mathrm{fid}}` (ref:512-580).  The peak positions come from the fiducial cosmology once (host);
    every call then fits and evaluates four natural cubic splines over all spectra at once on the device."""
    name = 'peakaverage_cuda'

    def _prepare(self):
        from scipy import signal
        from .utils import LeastSquareSolver
        index = np.flatnonzero((self.k >= 1e-3) & (self.k <= 1.))
        k_fid = self.k[index]
        ratio = _pk_lin(self.cosmo_fid, k_fid) / _pk_nowiggle(self.cosmo_fid, k_fid)                    # ref:527-533
        gradient = np.array([k_fid**(i - 1) for i in range(4)])
        constraint_gradient = np.column_stack([gradient[..., 0], gradient[..., 1] - gradient[..., 0], gradient[..., -1], gradient[..., -2] - gradient[..., -1]])
        solver = LeastSquareSolver(gradient, precision=k_fid**2, constraint_gradient=constraint_gradient, compute_inverse=False)
        solver(ratio, constraint=np.array([ratio[..., 0], ratio[..., 1] - ratio[..., 0], ratio[..., -1], ratio[..., -2] - ratio[..., -1]]))
        pknow_correction = solver.model()
        ik0 = np.searchsorted(k_fid, 1e-2, side='right') + 1
        self.k_peaks, self.pad_peaks = [], []
        for si in [1., -1.]:                                                                            # ref:541-549
            ik = signal.find_peaks(si * ratio[ik0:] / pknow_correction[ik0:])[0] + ik0
            npadlow = index[0]
            ik = ik + npadlow
            ikmax = max(index[-1], ik[-1] + 1)
            self.pad_peaks.append((int(npadlow), len(ik), int(self.k.size - ikmax)))
            self.k_peaks.append(self.k[np.concatenate([np.arange(npadlow), ik, np.arange(ikmax, self.k.size)], axis=0)])

    def _interp(self, xh, xl, x, y):
        from .interp import Interpolator1D
        logx = np.log10(x)
        interp = Interpolator1D(logx, y, k=3, extrap=True, assume_sorted=True)                          # ref:553
        toret = 0.
        for xx in [xh, xl]:
            logxx = np.log10(xx)
            toret = toret + Interpolator1D(logxx, interp(logxx), k=3, assume_sorted=True)(logx)         # ref:554-557
        return toret / 2.

    def _compute(self):
        rescale = self.rs_drag_ratio()
        rescale = [np.concatenate([np.linspace(1., rescale, npad[0]), np.full(npad[1], rescale), np.linspace(rescale, 1., npad[2])]) for npad in self.pad_peaks]
        pknow = _like(_pk_nowiggle(self.cosmo, self.k), self.pk)[:, None]                                # ref:564
        self.pknow = self._interp(self.k_peaks[0] / rescale[0], self.k_peaks[1] / rescale[1], self.k, self.pk / pknow) * pknow


_xi_registry = {}


class BaseCorrelationFunctionBAOFilter(_CosmoMixin):
    """Base BAO filter for the correlation function (ref:703-833)."""
    name = 'base'

    def __init__(self, xi_interpolator, cosmo=None, cosmo_fid=None, **kwargs):
        self._cosmo_fid = cosmo_fid
        self._cosmo = cosmo
        self.xi_interpolator = xi_interpolator
        self.set_s(**kwargs)
        self.set_xi(xi_interpolator, cosmo=cosmo)
        self._prepare()
        self._compute()
        self.xi, self.xinow = (x.reshape(self.shape) for x in (self.xi, self.xinow))

    def _prepare(self):
        """Anything that can be done once."""

    def set_s(self, ns=1024):
        self.s = np.geomspace(self.xi_interpolator.extrap_smin, self.xi_interpolator.extrap_smax, ns)   # ref:757

    def set_xi(self, xi_interpolator, cosmo=None):
        self._cosmo = cosmo                                                                             # ref:760 (the reference resets it here)
        self.xi_interpolator = xi_interpolator
        interp = xi_interpolator
        if hasattr(interp, 'z') and getattr(interp, 'z', None) is not None and np.ndim(interp.z) > 0 and hasattr(interp, 'growth_factor_sq'):
            self.xi = interp(self.s, interp.z, ignore_growth=True)
        else:
            self.xi = interp(self.s)
        self.shape = tuple(self.xi.shape)
        self.xi = self.xi.reshape(self.xi.shape[0], -1)

    def __call__(self, xi_interpolator, cosmo=None):
        self.set_xi(xi_interpolator, cosmo=cosmo)
        self._compute()
        self.xi, self.xinow = (x.reshape(self.shape) for x in (self.xi, self.xinow))
        return self

    def smooth_xi_interpolator(self, **kwargs):
        return self.xi_interpolator.clone(s=self.s, xi=self.xinow, **kwargs)

    def smooth_pk_interpolator(self, **kwargs):
        return self.smooth_xi_interpolator().to_pk(**kwargs)


class Kirkby2013CorrelationFunctionBAOFilter(BaseCorrelationFunctionBAOFilter):
    """Cut the BAO peak of the correlation function and bridge it with a polynomial s^1 .. s^-3 fitted on both sides (ref:835-909)."""
    name = 'kirkby2013_cuda'

    def __init__(self, xi_interpolator, srange_left=(50., 82.), srange_right=(150., 190.), rescale_sbox=True, cosmo=None, **kwargs):
        self.srange_left = np.asarray(srange_left, dtype='f8')
        self.srange_right = np.asarray(srange_right, dtype='f8')
        self.rescale_sbox = rescale_sbox
        super(Kirkby2013CorrelationFunctionBAOFilter, self).__init__(xi_interpolator, cosmo=cosmo, **kwargs)

    def _prepare(self):
        factor = 2.                                                                                     # ref:884-893
        self.smask = (self.s >= self.srange_left[0] / factor) & (self.s <= self.srange_right[1] * factor)
        self.model = np.array([self.s**(1 - i) for i in range(5)])
        frac = 1. / 100.
        shift_center = (self.srange_right[0] - self.srange_left[1]) * frac
        self.window = (np.concatenate([[self.srange_left[0] * (1. - frac)], self.srange_left,
                                       [self.srange_left[1] + shift_center, self.srange_right[0] - shift_center],
                                       self.srange_right, [self.srange_right[1] * (1. + frac)]], axis=0),
                       np.array([0., 1., 1., 0., 0., 1., 1., 0.]))

    def _compute(self):
        from .utils import LeastSquareSolver
        rescale = self.rs_drag_ratio() if self.rescale_sbox else 1.                                     # ref:897-899
        precision = np.interp(self.s[self.smask] / rescale, self.window[0], self.window[1], left=0., right=0.)
        center = np.interp(self.s / rescale, self.window[0][2:-2], 1. - self.window[1][2:-2], left=0., right=0.)
        solver = LeastSquareSolver(self.model[..., self.smask], precision=precision, compute_inverse=False)
        xi, was_numpy = _on_device(self.xi, getattr(self, '_device', None))
        lo, hi = int(np.flatnonzero(self.smask)[0]), int(np.flatnonzero(self.smask)[-1]) + 1
        params = solver(xi[lo:hi].T.contiguous())                                                       # ref:906
        model = params @ _like(self.model, xi)
        center = _like(center, xi)
        xinow = (xi.T * (1. - center) + model * center).T.contiguous()                                  # ref:908
        self.xinow = xinow.cpu().numpy() if was_numpy else xinow


_xi_registry[Kirkby2013CorrelationFunctionBAOFilter.name] = Kirkby2013CorrelationFunctionBAOFilter


def CorrelationFunctionBAOFilter(xi_interpolator, engine='kirkby2013_cuda', **kwargs):
    """Factory (ref:924-933); 'kirkby2013' is accepted as an alias."""
    name = engine.lower()
    if name == 'kirkby2013':
        name = 'kirkby2013_cuda'
    try:
        cls = _xi_registry[name]
    except KeyError:
        raise ValueError('Correlation function BAO filter {} is unknown; cosmoprimo_b200 provides {}'.format(engine, sorted(_xi_registry)))
    return cls(xi_interpolator, **kwargs)


def PowerSpectrumBAOFilter(pk_interpolator, engine='wallish2018_cuda', **kwargs):
    """Factory (ref:912-921): ``engine`` is one of the registered names; 'wallish2018' is accepted as an alias."""
    name = engine.lower()
    if name in ('wallish2018', 'ehpoly', 'peakaverage'):
        name = name + '_cuda'
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
    for cls in _registry.values():
        ref.RegisteredPowerSpectrumBAOFilter._registry[cls.name] = cls
    for cls in _xi_registry.values():
        ref.RegisteredCorrelationFunctionBAOFilter._registry[cls.name] = cls
    return ref
