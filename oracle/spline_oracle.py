"""
ORACLE (test infrastructure, not product code): CPU restatement of cosmoprimo's 1-D cubic-spline path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

* :func:`interpolator1d` restates ``cosmoprimo/jax.py:134-196`` (``Interpolator1D``, numpy path) on top of the same
  third-party call the reference makes, ``scipy.interpolate.CubicSpline`` (scipy is an unpinned dependency of the
  reference, pyproject.toml:12; 1.18.1 here), so it is bit-identical to the reference on one machine.
* :func:`cubic_spline_slopes` / :func:`cubic_spline_eval` restate what scipy does inside (``_cubic.py``: slope system,
  banded solve; ``PPoly`` power basis) without calling it, and are checked against it in tests/test_spline_oracle.py.

Parity pinned by golden vectors generated from the reference's own ``Interpolator1D`` (tools/make_golden.py).
"""

import numpy as np


def cubic_spline_slopes(x, y, bc='natural'):
    """
    Knot slopes of the C2 cubic spline through (x, y[:, col]) for every column; ``bc`` 'natural' (y''=0 at both ends)
    or 'clamped' (y'=0 at both ends).  Row i: dx_i s_{i-1} + 2 (dx_{i-1}+dx_i) s_i + dx_{i-1} s_{i+1}
    = 3 (dx_i m_{i-1} + dx_{i-1} m_i); Thomas elimination (the matrix is strictly diagonally dominant by rows).
    """
    x = np.asarray(x, dtype='f8')
    y = np.asarray(y, dtype='f8')
    flat = y.reshape(x.size, -1)
    n = x.size
    dx = np.diff(x)
    m = np.diff(flat, axis=0) / dx[:, None]
    lo, di, up = np.zeros(n), np.zeros(n), np.zeros(n)
    rhs = np.zeros_like(flat)
    lo[1:-1], di[1:-1], up[1:-1] = dx[1:], 2 * (dx[:-1] + dx[1:]), dx[:-1]
    rhs[1:-1] = 3 * (dx[1:, None] * m[:-1] + dx[:-1, None] * m[1:])
    if bc == 'clamped':
        di[0] = di[-1] = 1.
    else:
        di[0], up[0], rhs[0] = 2 * dx[0], dx[0], 3 * (flat[1] - flat[0])
        di[-1], lo[-1], rhs[-1] = 2 * dx[-1], dx[-1], 3 * (flat[-1] - flat[-2])
    cp = np.zeros(n)
    dp = np.zeros_like(flat)
    cp[0] = up[0] / di[0]
    dp[0] = rhs[0] / di[0]
    for i in range(1, n):
        w = 1. / (di[i] - lo[i] * cp[i - 1])
        cp[i] = up[i] * w
        dp[i] = (rhs[i] - lo[i] * dp[i - 1]) * w
    s = np.empty_like(flat)
    s[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        s[i] = dp[i] - cp[i] * s[i + 1]
    return s.reshape(y.shape)


def cubic_spline_eval(x, y, s, xq, nu=0, extrapolate=False):
    """nu-th derivative at ``xq`` of the piecewise cubic with knot values ``y`` and slopes ``s`` (scipy's PPoly basis)."""
    x = np.asarray(x, dtype='f8')
    flat, sl = np.asarray(y, dtype='f8').reshape(x.size, -1), np.asarray(s, dtype='f8').reshape(x.size, -1)
    xq = np.asarray(xq, dtype='f8')
    i = np.clip(np.searchsorted(x, xq, side='right') - 1, 0, x.size - 2)
    dx = (x[i + 1] - x[i])[:, None]
    m = (flat[i + 1] - flat[i]) / dx
    t = (sl[i] + sl[i + 1] - 2 * m) / dx
    c0, c1, c2, c3 = t / dx, (m - sl[i]) / dx - t, sl[i], flat[i]
    d = (xq - x[i])[:, None]
    if nu == 0: out = c3 + d * (c2 + d * (c1 + d * c0))
    elif nu == 1: out = c2 + d * (2 * c1 + d * 3 * c0)
    elif nu == 2: out = 2 * c1 + 6 * c0 * d
    else: out = 6 * c0 + 0 * d
    if not extrapolate:
        out = np.where(((xq >= x[0]) & (xq <= x[-1]))[:, None], out, np.nan)
    return out.reshape(xq.shape + np.shape(y)[1:])


def interpolator1d(x, fun, interp_x='lin', interp_fun='lin', extrap=False, assume_sorted=False):
    """Restatement of Interpolator1D (jax.py:139-196), k=3, numpy path.  Returns a callable ``(xq, dx=0) -> values``."""
    from scipy.interpolate import CubicSpline
    x = np.array(x, dtype='f8')
    fun = np.array(fun, dtype='f8')
    shape = fun.shape[1:]
    if not assume_sorted:                                         # :147-149
        ix = np.argsort(x)
        x, fun = x[ix], fun[ix]
    xmin, xmax = x[0], x[-1]                                      # :150
    if interp_x == 'log': x = np.log10(x)                         # :152
    if interp_fun == 'log': fun = np.log10(fun)                   # :153
    fun = fun.reshape(x.size, -1)
    mask_nan = ~np.isnan(fun).all(axis=0)                         # :161
    fun = fun[..., mask_nan]
    spline = None
    if fun.size and not np.isnan(fun).any():                      # :168-172
        spline = CubicSpline(x, fun, axis=0, bc_type='natural', extrapolate=bool(extrap))

    def call(xq, dx=0):
        dtype = np.result_type(*[a.dtype for a in (xq,) if hasattr(a, 'dtype')] or [np.float64])   # utils.py:88-95
        if not np.issubdtype(dtype, np.floating): dtype = np.float64
        xq = np.asarray(xq, dtype=dtype)
        toret_shape = xq.shape + shape
        xq = xq.ravel()
        mask_x = (xq >= xmin) & (xq <= xmax)                      # :188
        if interp_x == 'log': xq = np.log10(xq)
        tmp = spline(xq, nu=dx) if spline is not None else np.full(xq.shape + fun.shape[1:], np.nan)
        if interp_fun == 'log': tmp = 10**tmp                     # :191
        if not extrap: tmp = np.where(mask_x, tmp.T, np.nan).T    # :192
        toret = np.full((xq.size, mask_nan.size), np.nan)         # :194-195
        toret[..., mask_nan] = tmp
        return toret.astype(dtype).reshape(toret_shape)

    return call


def interpolator2d(x, y, fun, interp_x='lin', interp_y='lin', interp_fun='lin', extrap=False):
    """Restatement of Interpolator2D (jax.py:213-277), kx = ky = 3, numpy path: the same third-party call,
    ``scipy.interpolate.RectBivariateSpline(x, y, fun, s=0)`` (FITPACK).  Returns ``(xq, yq, grid=True) -> values``."""
    from scipy.interpolate import RectBivariateSpline
    x, y, fun = (np.array(a, dtype='f8') for a in (x, y, fun))
    ix, iy = np.argsort(x), np.argsort(y)                                       # :224-226
    x, y, fun = x[ix], y[iy], fun[np.ix_(ix, iy)]
    xmin, xmax, ymin, ymax = x[0], x[-1], y[0], y[-1]
    if interp_x == 'log': x = np.log10(x)
    if interp_y == 'log': y = np.log10(y)
    if interp_fun == 'log': fun = np.log10(fun)
    spline = RectBivariateSpline(x, y, fun, kx=3, ky=3, s=0)                    # :242-243

    def call(xq, yq, grid=True):
        xq, yq = np.asarray(xq, dtype='f8'), np.asarray(yq, dtype='f8')
        shape = xq.shape + yq.shape if grid else xq.shape
        xq, yq = xq.ravel(), yq.ravel()
        mx, my = (xq >= xmin) & (xq <= xmax), (yq >= ymin) & (yq <= ymax)       # :254-256
        mask = mx[:, None] & my if grid else mx & my
        if interp_x == 'log': xq = np.log10(xq)
        if interp_y == 'log': yq = np.log10(yq)
        if grid:
            i_x, i_y = np.argsort(xq), np.argsort(yq)                           # :268-270
            tmp = spline(xq[i_x], yq[i_y], grid=True)[np.ix_(np.argsort(i_x), np.argsort(i_y))]
        else:
            tmp = spline(xq, yq, grid=False)
        if interp_fun == 'log': tmp = 10**tmp
        return (tmp if extrap else np.where(mask, tmp, np.nan)).reshape(shape)

    return call


def interpolator2d_factorised(x, y, fun, xq, yq):
    """The formulation the CUDA path uses (cosmoprimo_b200/interp.py::Interpolator2D): FITPACK's interpolating bicubic
    spline has not-a-knot ends, and tensor-product interpolation factorises into 1-D not-a-knot splines along y for every
    x knot followed by 1-D not-a-knot splines along x.  Library-free of FITPACK; checked against :func:`interpolator2d`."""
    from scipy.interpolate import CubicSpline
    vals = CubicSpline(y, np.asarray(fun, dtype='f8').T, axis=0, bc_type='not-a-knot')(yq)       # (nyq, nx)
    return CubicSpline(x, vals.T, axis=0, bc_type='not-a-knot')(xq)                               # (nxq, nyq)
