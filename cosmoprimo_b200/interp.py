"""
Batched 1-D cubic-spline interpolation on the GPU: drop-in for the numpy path of ``cosmoprimo.jax.Interpolator1D``
(``cosmoprimo/jax.py:134-196``, cited as ``ref:LINE``): natural cubic spline along axis 0 of ``fun``, optional log10
abscissa / ordinate, NaN outside the fitted range unless ``extrap``, all-NaN columns passed through, float32 output
only for float32 queries.

Fit (tridiagonal slope system, one thread per column) and evaluation run in ``csrc/cpf_spline.cu`` behind
``cpf_spline_create`` / ``cpf_spline_eval``; the fitted spline lives on the device.  numpy in -> numpy out; CUDA
arrays (torch / ``__cuda_array_interface__`` / DLPack) in -> torch tensors out.
"""

import ctypes
import threading
import collections

import numpy as np

from . import _lib
from . import _buffers as _buf


def _bcast_dtype(*args):
    """float32 only if every array argument is float32, else float64 (``cosmoprimo/utils.py:88-95``)."""
    dtypes = [np.dtype(str(a.dtype).replace('torch.', '')) for a in args if hasattr(a, 'dtype')]
    if not dtypes:
        return np.dtype('f8')
    out = np.result_type(*dtypes)
    return out if np.issubdtype(out, np.floating) else np.dtype('f8')


_BC_CODES = {'natural': 0, 'clamped': 1, 'not-a-knot': 2}


class _DeviceSpline(object):
    """Owner of a ``cpf_spline*``."""

    def __init__(self, xbuf, ybuf, nx, ncols, bc, log_x, log_y, extrap, device, stream):
        handle = ctypes.c_void_p()
        rc = _lib.load().cpf_spline_create(ctypes.byref(handle), xbuf.ptr, ybuf.ptr, nx, ncols, bc, int(log_x), int(log_y), int(extrap),
                                           int(ybuf.on_device), device, stream)
        _lib.check(rc)
        self.handle, self.device = handle, device

    def __del__(self):
        handle, self.handle = getattr(self, 'handle', None), None
        if handle:
            try:
                _lib.load().cpf_spline_destroy(handle)
            except Exception:
                pass


class Interpolator1D(object):
    """
    1-D cubic-spline interpolation along axis 0 of ``fun`` on a CUDA device.

    Parameters
    ----------
    x : array (nx,)
        Abscissae (sorted unless ``assume_sorted`` is False, in which case they are sorted here, ref:147-149).
    fun : array (nx, ...)
        Ordinates; numpy array or CUDA array.
    k : int, default=3
        Spline order; only the cubic spline of the hot path is implemented.
    interp_x, interp_fun : 'lin' or 'log'
        Interpolate in log10 of the abscissa / ordinate (ref:152-153, 189-191).
    extrap : bool, default=False
        If False, NaN outside ``[x.min(), x.max()]``.
    bc_type : 'natural' (the reference's choice, ref:172), 'clamped' (used by the Wallish2018 filter) or 'not-a-knot'
        (the ends of FITPACK's interpolating splines: building block of :class:`Interpolator2D`).
    device : int, default=None
        CUDA device for host input.
    """

    def __init__(self, x, fun, k=3, interp_x='lin', interp_fun='lin', extrap=False, assume_sorted=False, bc_type='natural', device=None):
        if k != 3:
            raise NotImplementedError('cosmoprimo_b200.Interpolator1D implements the cubic spline (k=3) only')
        if bc_type not in _BC_CODES:
            raise ValueError('bc_type must be one of {}'.format(sorted(_BC_CODES)))
        _lib.load()
        self.interp_x, self.interp_fun = str(interp_x), str(interp_fun)
        self.extrap = bool(extrap)
        x = np.array(x, dtype='f8').ravel()
        on_device = _buf.is_device_array(fun)
        if on_device:
            ybuf = _buf.as_input(fun, dtype='f8')
            fun_t = ybuf.obj
        else:
            fun_t = np.array(fun, dtype='f8')
        if fun_t.shape[0] != x.size:
            raise ValueError('fun has {} samples along axis 0, x has {}'.format(fun_t.shape[0], x.size))
        self.shape = tuple(fun_t.shape[1:])
        if not assume_sorted:
            ix = np.argsort(x)
            if not np.array_equal(ix, np.arange(x.size)):
                x = x[ix]
                fun_t = fun_t[ix] if not on_device else fun_t[_buf._torch().as_tensor(ix, device=fun_t.device)]
        self.xmin, self.xmax = x[0], x[-1]
        flat = fun_t.reshape(x.size, -1)
        # all-NaN columns are passed through as NaN, any other NaN poisons the whole fit (ref:161-172)
        if on_device:
            flat = flat.contiguous()
            flags = np.zeros(max(int(flat.shape[1]), 1), dtype='u1')                 # one read of the table (cpf_column_nan_flags)
            _lib.check(_lib.load().cpf_column_nan_flags(flat.data_ptr(), x.size, int(flat.shape[1]), int(self.interp_fun == 'log'), flags.ctypes.data,
                                                        flat.device.index, _buf.current_stream(flat.device.index)))
            flags = flags[:int(flat.shape[1])]
            self._mask_nan = flags != 1
            poisoned = bool((flags == 2).any())
        else:
            with np.errstate(invalid='ignore'):
                isnan = np.isnan(flat) | ((flat < 0) if self.interp_fun == 'log' else False)
            self._mask_nan = ~isnan.all(axis=0)
            poisoned = bool(isnan[:, self._mask_nan].any())
        self._ncols = int(flat.shape[1])
        self._spline = None
        self._on_device = on_device
        if x.size < 2:
            raise ValueError('need at least two knots')
        if self._mask_nan.any() and not poisoned:
            _lib.require_device()
            ybuf = _buf.as_input(flat, dtype='f8')
            dev = ybuf.device if ybuf.on_device else (device if device is not None else _buf.default_device())
            stream = _buf.current_stream(dev) if ybuf.on_device else None
            if ybuf.on_device:
                xbuf = _buf.as_input(_buf._torch().as_tensor(x, device=ybuf.obj.device), dtype='f8')
            else:
                xbuf = _buf.as_input(x, dtype='f8')
            # NaN columns simply propagate NaN through their own (independent) solve: no need to drop them
            self._spline = _DeviceSpline(xbuf, ybuf, x.size, self._ncols, _BC_CODES[bc_type], self.interp_x == 'log',
                                         self.interp_fun == 'log', self.extrap, dev, stream)
            self._device = dev

    @classmethod
    def padlog(cls, x_padded, fun, extrap=False, device=None):
        """
        The interpolator ``PowerSpectrumInterpolator1D(extrap_pk='log')`` builds (``cosmoprimo/interpolator.py:42-87, 343-351``):
        natural spline of log10(fun) in log10(x) with two power-law continuation knots on each side, in one pass over ``fun``
        (``cpf_spline_create_padlog``: logarithms, continuation rows, NaN screening and the fit, no intermediate array).
        ``x_padded`` (nx + 4,) are the padded knots, ``fun`` (nx, ...) the tabulated values, numpy or CUDA array.
        """
        lib = _lib.load()
        _lib.require_device()
        self = cls.__new__(cls)
        self.interp_x = self.interp_fun = 'log'
        self.extrap = bool(extrap)
        x = np.ascontiguousarray(x_padded, dtype='f8').ravel()
        ybuf = _buf.as_input(fun, dtype='f8')
        nx = x.size - 4
        if nx < 2 or ybuf.shape[0] != nx:
            raise ValueError('fun has {} samples along axis 0, x_padded has {} knots (4 of them continuation knots)'.format(ybuf.shape[0], x.size))
        self.shape = tuple(ybuf.shape[1:])
        self.xmin, self.xmax = x[0], x[-1]
        self._ncols = int(np.prod(self.shape, dtype='i8'))
        self._on_device = ybuf.on_device
        self._spline = None
        dev = ybuf.device if ybuf.on_device else (device if device is not None else _buf.default_device())
        stream = _buf.current_stream(dev) if ybuf.on_device else None
        flags = np.zeros(max(self._ncols, 1), dtype='u1')
        handle = ctypes.c_void_p()
        _lib.check(lib.cpf_spline_create_padlog(ctypes.byref(handle), x.ctypes.data, ybuf.ptr, nx, self._ncols, int(self.extrap),
                                                flags.ctypes.data, int(ybuf.on_device), dev, stream))
        spline = _DeviceSpline.__new__(_DeviceSpline)
        spline.handle, spline.device = handle, dev
        flags = flags[:self._ncols]
        # all-NaN columns are passed through as NaN, any other NaN poisons the whole fit (ref:161-172)
        self._mask_nan = flags != 1
        if self._mask_nan.any() and not (flags == 2).any():
            self._spline, self._device = spline, dev
        return self

    def __call__(self, x, bounds_error=False, dx=0):
        """Evaluate the spline (or its ``dx``-th derivative) at ``x``; result has shape ``x.shape + fun.shape[1:]``."""
        dtype = _bcast_dtype(x)
        q_on_device = _buf.is_device_array(x)
        if q_on_device:
            qbuf = _buf.as_input(x, dtype='f8')
            q = qbuf.obj.reshape(-1)
            qshape = tuple(qbuf.shape)
        else:
            q = np.asarray(x, dtype=dtype).astype('f8').ravel()
            qshape = np.shape(x)
        if bounds_error:
            qh = q.cpu().numpy() if q_on_device else q
            if qh.size and not ((qh >= self.xmin) & (qh <= self.xmax)).all():
                raise ValueError('input outside of extrapolation range (min: {} vs. {}; max: {} vs. {})'.format(qh.min(), self.xmin, qh.max(), self.xmax))
        out_shape = qshape + self.shape
        nq = int(np.prod(qshape, dtype='i8'))
        want_device = self._on_device or q_on_device
        if self._spline is None or nq == 0 or self._ncols == 0:
            if want_device:
                torch = _buf._torch()
                dev = q.device if q_on_device else 'cuda'
                return torch.full(out_shape, float('nan'), dtype=torch.float32 if dtype == np.float32 else torch.float64, device=dev)
            return np.full(out_shape, np.nan, dtype=dtype)
        lib = _lib.load()
        dev = self._spline.device
        if want_device:
            torch = _buf._torch()
            if not q_on_device:
                q = _device_copy(np.ascontiguousarray(q), dev)          # cached by content: the same grids come back call after call
            qbuf = _buf.as_input(q.contiguous(), dtype='f8')
            out = _buf.empty_like_kind(qbuf, (nq, self._ncols), dtype='f8')
            stream = _buf.current_stream(dev)
        else:
            qbuf = _buf.as_input(q, dtype='f8')
            out = _buf.empty_like_kind(qbuf, (nq, self._ncols), dtype='f8')
            stream = None
        # evaluate in slabs of 65535 queries (grid.y limit of the kernel)
        step = 65535
        for start in range(0, nq, step):
            cnt = min(step, nq - start)
            rc = lib.cpf_spline_eval(self._spline.handle, qbuf.ptr + 8 * start, cnt, int(dx), out.ptr + 8 * start * self._ncols,
                                     int(qbuf.on_device), stream)
            _lib.check(rc)
        res = out.obj
        if want_device:
            if dtype == np.float32:
                res = res.to(_buf._torch().float32)
            return res.reshape(out_shape)
        return res.astype(dtype, copy=False).reshape(out_shape)

    def eval_rows(self, x, dx=0):
        """
        Same values as ``self(x)`` for 1-D ``x``, transposed: (ncols, nq) float64 with one row per spline -- the layout the
        FFTLog entry points read -- written directly by the kernel (``cpf_spline_eval_t``), no transposition pass.
        Result kind (numpy / torch) follows the fitted table.
        """
        q = np.asarray(x, dtype='f8').ravel()
        nq = q.size
        if self._spline is None or nq == 0 or self._ncols == 0:
            out = np.full((self._ncols, nq), np.nan)
            return _buf._torch().as_tensor(out, device='cuda') if self._on_device else out
        lib = _lib.load()
        dev = self._spline.device
        if self._on_device:
            torch = _buf._torch()
            qbuf = _buf.as_input(_device_copy(q, dev), dtype='f8')      # the wavenumber grid of sigma_r / to_xi repeats across calls
            stream = _buf.current_stream(dev)
        else:
            qbuf = _buf.as_input(q, dtype='f8')
            stream = None
        out = _buf.empty_like_kind(qbuf, (self._ncols, nq), dtype='f8')
        _lib.check(lib.cpf_spline_eval_t(self._spline.handle, qbuf.ptr, nq, int(dx), out.ptr, int(qbuf.on_device), stream))
        return out.obj


class Interpolator2D(object):
    """
    Bicubic interpolation on a rectangular grid, drop-in for the numpy path of ``cosmoprimo.jax.Interpolator2D``
    (ref:213-277), which wraps ``scipy.interpolate.RectBivariateSpline(x, y, fun, kx=3, ky=3, s=0)``: FITPACK's interpolating
    tensor-product spline has not-a-knot ends in both directions, and tensor-product interpolation factorises, so the value
    at (xq, yq) is obtained by 1-D not-a-knot splines along y for every x knot (fitted once, here), then a 1-D not-a-knot
    spline along x through the values at yq (fitted per call: nx knots x nyq columns).  Both steps are the batched spline
    kernels of ``csrc/cpf_spline.cu``.  ``fun`` has shape (nx, ny); numpy or CUDA array.
    """

    def __init__(self, x, y, fun, kx=3, ky=3, interp_x='lin', interp_y='lin', interp_fun='lin', extrap=False, assume_sorted=False, device=None):
        if kx != 3 or ky != 3:
            raise NotImplementedError('cosmoprimo_b200.Interpolator2D implements bicubic splines (kx = ky = 3) only')
        self.interp_x, self.interp_y, self.interp_fun = str(interp_x), str(interp_y), str(interp_fun)
        self.extrap = bool(extrap)
        x, y = (np.array(xx, dtype='f8').ravel() for xx in (x, y))
        on_device = _buf.is_device_array(fun)
        fun = _buf.as_input(fun, dtype='f8').obj if on_device else np.array(fun, dtype='f8')
        if tuple(fun.shape) != (x.size, y.size):
            raise ValueError('fun must have shape ({}, {}), got {}'.format(x.size, y.size, tuple(fun.shape)))
        if x.size < 4 or y.size < 4:
            raise NotImplementedError('bicubic interpolation needs at least 4 knots in each direction')
        if not assume_sorted:
            ix, iy = np.argsort(x), np.argsort(y)
            x, y = x[ix], y[iy]
            if on_device:
                torch = _buf._torch()
                fun = fun[torch.as_tensor(ix, device=fun.device)][:, torch.as_tensor(iy, device=fun.device)]
            else:
                fun = fun[np.ix_(ix, iy)]
        self.xmin, self.xmax, self.ymin, self.ymax = x[0], x[-1], y[0], y[-1]
        self._on_device = on_device
        self._device = device
        self._xt = np.log10(x) if self.interp_x == 'log' else x
        yt = np.log10(y) if self.interp_y == 'log' else y
        if self.interp_fun == 'log':
            fun = _buf._torch().log10(fun) if on_device else np.log10(fun)
        funT = fun.T.contiguous() if on_device else np.ascontiguousarray(fun.T)       # (ny, nx): knots of the y splines along axis 0
        self._along_y = Interpolator1D(yt, funT, bc_type='not-a-knot', extrap=True, assume_sorted=True, device=device)

    def __call__(self, x, y, grid=True, bounds_error=False, rows=False):
        """``rows=True`` (grid only, 1-D queries): the float64 result transposed, (ny, nx), one row per y query."""
        dtype = _bcast_dtype(x, y)
        x, y = (np.asarray(xx, dtype=dtype).astype('f8') for xx in (x, y))
        shape = x.shape + y.shape if grid else x.shape
        x, y = x.ravel(), y.ravel()
        masks = []
        for q, lo, hi in ((x, self.xmin, self.xmax), (y, self.ymin, self.ymax)):
            m = (q >= lo) & (q <= hi)
            if bounds_error and not m.all():
                raise ValueError('input outside of extrapolation range (min: {} vs. {}; max: {} vs. {})'.format(q.min(), lo, q.max(), hi))
            masks.append(m)
        mask = masks[0][:, None] & masks[1] if grid else masks[0] & masks[1]
        # FITPACK evaluates queries outside the table at the nearest edge (bispev clamps its arguments): with ``extrap`` the
        # reference therefore returns edge values, not extended polynomials
        xt = np.log10(np.clip(x, self.xmin, self.xmax)) if self.interp_x == 'log' else np.clip(x, self.xmin, self.xmax)
        yt = np.log10(np.clip(y, self.ymin, self.ymax)) if self.interp_y == 'log' else np.clip(y, self.ymin, self.ymax)
        if x.size == 0 or y.size == 0:
            out = np.zeros(shape, dtype=dtype)
            return _buf._torch().as_tensor(out, device='cuda') if self._on_device else out
        vals = self._along_y(yt)                                              # (nyq, nx)
        valsT = vals.T.contiguous() if self._on_device else np.ascontiguousarray(vals.T)      # (nx, nyq)
        if grid and rows:
            tmp = Interpolator1D(self._xt, valsT, bc_type='not-a-knot', extrap=True, assume_sorted=True, device=self._device).eval_rows(xt)   # (nyq, nxq)
            mask, shape, dtype = mask.T, (y.size, x.size), np.dtype('f8')
        elif grid:
            tmp = Interpolator1D(self._xt, valsT, bc_type='not-a-knot', extrap=True, assume_sorted=True, device=self._device)(xt)   # (nxq, nyq)
        else:
            # pairs (x_i, y_i): column i of the x splines evaluated at x_i only
            step, parts = 2048, []
            for start in range(0, x.size, step):
                sl = slice(start, start + step)
                full = Interpolator1D(self._xt, valsT[:, sl], bc_type='not-a-knot', extrap=True, assume_sorted=True, device=self._device)(xt[sl])
                parts.append(full.diagonal() if self._on_device else np.diagonal(full))
            tmp = (_buf._torch().cat(parts) if self._on_device else np.concatenate(parts))
        if self.interp_fun == 'log':
            tmp = 10**tmp
        if self._on_device:
            torch = _buf._torch()
            if not self.extrap:
                tmp = torch.where(torch.as_tensor(mask, device=tmp.device), tmp, torch.full_like(tmp, float('nan')))
            return tmp.to(torch.float32 if dtype == np.float32 else torch.float64).reshape(shape)
        if not self.extrap:
            tmp = np.where(mask, tmp, np.nan)
        return tmp.astype(dtype).reshape(shape)


_DEVICE_COPIES = collections.OrderedDict()       # LRU, guarded by a lock like the plan caches of fftlog.py (calls may come from several threads)
_DEVICE_COPIES_LOCK = threading.Lock()


def _device_copy(a, dev):
    """Device copy of a small host grid, cached by content: the same knots / radii come back on every sigma(r) call and a
    pageable host-to-device copy synchronises the stream."""
    torch = _buf._torch()
    if a.nbytes > 65536:                       # large query sets are not worth remembering
        return torch.as_tensor(a, device=torch.device('cuda', dev))
    key = (dev, a.size, a.tobytes())
    with _DEVICE_COPIES_LOCK:
        hit = _DEVICE_COPIES.get(key)
        if hit is not None:
            _DEVICE_COPIES.move_to_end(key)
            return hit
    new = torch.as_tensor(a, device=torch.device('cuda', dev))
    with _DEVICE_COPIES_LOCK:
        hit = _DEVICE_COPIES.setdefault(key, new)
        while len(_DEVICE_COPIES) > 16:
            _DEVICE_COPIES.popitem(last=False)
    return hit


def spline_eval_rows(x, fun, xq, bc_type='natural', window=64, extrap=False, device=None, sqrt=False):
    """
    Cubic splines along the LAST axis of ``fun`` (rows, nx) -- the layout FFTLog returns -- on the shared knots ``x``
    (nx,), evaluated at ``xq`` (nq,): returns (nq, rows), i.e. what the reference obtains with
    ``Interpolator1D(x, fun.T, assume_sorted=True)(xq)`` (``cosmoprimo/interpolator.py:289``) without the two transposes
    and without a global fit (``cpf_spline_eval_rows``: the value is a weighted sum of the ordinates; the slope system is
    solved on ``window`` knots either side of the bracketing interval, ``window=0``: all knots).  ``sqrt``: the kernel writes the
    square root of the value (sigma from the variance) instead of a second pass over the result.
    """
    if bc_type not in ('natural', 'clamped'):
        raise ValueError('bc_type must be "natural" or "clamped"')
    lib = _lib.load()
    _lib.require_device()
    x = np.ascontiguousarray(x, dtype='f8').ravel()
    q = np.ascontiguousarray(xq, dtype='f8').ravel()
    ybuf = _buf.as_input(fun, dtype='f8')
    if len(ybuf.shape) != 2 or ybuf.shape[1] != x.size:
        raise ValueError('fun must have shape (rows, {}), got {}'.format(x.size, tuple(ybuf.shape)))
    rows = int(ybuf.shape[0])
    if ybuf.on_device:
        dev = ybuf.device
        xbuf = _buf.as_input(_device_copy(x, dev))
        qbuf = _buf.as_input(_device_copy(q, dev))
        stream = _buf.current_stream(dev)
    else:
        dev = device if device is not None else _buf.default_device()
        xbuf, qbuf, stream = _buf.as_input(x), _buf.as_input(q), None
    out = _buf.empty_like_kind(ybuf, (q.size, rows))
    _lib.check(lib.cpf_spline_eval_rows(xbuf.ptr, ybuf.ptr, x.size, rows, qbuf.ptr, q.size, 1 if bc_type == 'clamped' else 0, int(window),
                                        int(bool(extrap)) | (2 if sqrt else 0), out.ptr, int(ybuf.on_device), dev, stream))
    return out.obj
