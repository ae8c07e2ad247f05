"""
ORACLE (test infrastructure, not product code): CPU restatement of cosmoprimo's Wallish2018 no-wiggle filter,
``Wallish2018PowerSpectrumBAOFilter._compute`` (cosmoprimo/bao_filter.py:361-431).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

The reference delegates the arithmetic to scipy (``scipy.fftpack.dst/idst`` -> ducc, ``CubicSpline`` -> LAPACK banded
solve); :func:`wallish2018` calls the same scipy entry points in the same order, so it is bit-identical to the
reference on one machine.  The reference has NO golden values for this filter ("parity unpinned" in its
own tests); parity is pinned here by vectors generated from the reference class itself (tools/make_golden.py ->
tests/golden/wallish_golden.npz).

Input is what the reference evaluates from its interpolator: ``pklin = pk_interpolator(klin)`` on
``klin = linspace(extrap_kmin, 2, 4096)`` (:364-369) and ``pkout = pk_interpolator(kout)`` on ``kout = self.k`` (:90-102).
"""

import numpy as np

MARGIN_FIRST, MARGIN_SECOND, OFFSET = 20, 5, (-10, 20)      # bao_filter.py:387-389


def find_boxes(dd):
    """(ibox0, ibox1) of one column of second derivatives (bao_filter.py:392-395); numpy argmax = first maximum."""
    a = dd[MARGIN_FIRST:-MARGIN_FIRST].argmax() + MARGIN_FIRST
    b = a + MARGIN_SECOND + dd[a + MARGIN_SECOND:-MARGIN_FIRST].argmax()
    return a + OFFSET[0], b + OFFSET[1]


def tophat(k, kmax=1., scale=1.):                            # bao_filter.py:425-431
    out = np.ones_like(k)
    mask = k > kmax
    out[mask] *= np.exp(-scale**2 * (k[mask] / kmax - 1.)**2)
    return out


def wallish2018(klin, pklin, kout, pkout, return_debug=False):
    """pknow [nk, ncols] (bao_filter.py:371-423)."""
    from scipy import fftpack, interpolate
    klin, kout = np.asarray(klin, dtype='f8'), np.asarray(kout, dtype='f8')
    pklin = np.asarray(pklin, dtype='f8').reshape(klin.size, -1)
    pkout = np.asarray(pkout, dtype='f8').reshape(kout.size, -1)
    kpk = np.log(klin[:, None] * pklin)                                                       # :371
    ffted = fftpack.dst(kpk, type=2, axis=0, norm='ortho', overwrite_x=False)                 # :372
    even, odd = ffted[::2].copy(), ffted[1::2].copy()                                         # :373-374
    xe, xo = 1 + np.arange(even.shape[0]), 1 + np.arange(odd.shape[0])                        # :376
    dd_even = interpolate.CubicSpline(xe, even, axis=0, bc_type='clamped', extrapolate=False)(xe, nu=2)   # :377-379
    dd_odd = interpolate.CubicSpline(xo, odd, axis=0, bc_type='clamped', extrapolate=False)(xo, nu=2)     # :380-382
    debug = dict(even=even.copy(), odd=odd.copy(), dd_even=dd_even, dd_odd=dd_odd)
    boxes = np.zeros((pklin.shape[1], 4), dtype='i4')
    for ic in range(pklin.shape[1]):                                                          # :404-405
        for half, (x, arr, dd) in enumerate([(xe, even, dd_even), (xo, odd, dd_odd)]):
            b0, b1 = find_boxes(dd[:, ic])
            boxes[ic, 2 * half:2 * half + 2] = b0, b1
            mask = np.ones(x.size, dtype=bool)
            mask[b0:b1 + 1] = False                                                           # :396-399
            spl = interpolate.CubicSpline(x[mask], arr[mask, ic] * x[mask]**2, axis=-1, bc_type='clamped', extrapolate=False)
            arr[:, ic] = spl(x) / x**2                                                        # :400-402
    merged = np.empty_like(ffted)                                                             # :409-411
    merged[::2], merged[1::2] = even, odd
    kpknow = fftpack.idst(merged, type=2, axis=0, norm='ortho', overwrite_x=False)            # :412
    pknow = np.exp(kpknow) / klin[:, None]                                                    # :413
    mask = (klin > 1e-2) & (klin < 1.5)                                                       # :415
    left, right = kout < 5e-4, kout > 2.                                                      # :417
    knots = np.concatenate([kout[left], klin[mask], kout[right]], axis=0)                     # :418
    vals = np.concatenate([pkout[left], pknow[mask], pkout[right]], axis=0)                   # :419
    pknow = interpolate.CubicSpline(knots, vals, axis=0, bc_type='clamped', extrapolate=False)(kout)   # :420
    th = tophat(kout, kmax=1., scale=20.)[:, None]                                            # :421
    wiggles = (pkout / pknow - 1.) * th + 1.                                                  # :422
    out = pkout / wiggles                                                                     # :423
    if return_debug:
        debug.update(even_now=even, odd_now=odd, boxes=boxes, knots=knots, kpknow=kpknow)
        return out, debug
    return out
