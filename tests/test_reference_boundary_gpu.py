"""The drop-in boundary exercised from the REFERENCE's side, on the GPU box: the unmodified reference, shipped as
baseline/_ref, is handed our engine / filter through the two zero-patch routes it already has:

* ``cosmoprimo.fftlog.get_fft_engine`` returns a non-string engine unchanged (ref fftlog.py:663);
* ``cosmoprimo.bao_filter.PowerSpectrumBAOFilter(..., engine=name)`` looks ``name`` up in the metaclass registry (ref bao_filter.py:22-31, 912-921).

Wallish2018 parity at scale is counted here too: every column whose four box indices differ from the reference's is reported (expected: none).
"""
import numpy as np
import pytest

from conftest import reference
from cosmoprimo_b200 import fftlog as F, synthetic as S
from cosmoprimo_b200 import bao_filter as B

pytestmark = pytest.mark.gpu


def lhs_pk(ncosmo, n, seed=5):
    k = np.geomspace(1e-5, 1e2, n)
    return k, S.eh_pk(k, S.lhs_cosmologies(ncosmo, seed=seed))


@pytest.mark.parametrize('cls,kw', [('PowerToCorrelation', {'ell': 0}), ('PowerToCorrelation', {'ell': [0, 2, 4]}), ('TophatVariance', {}),
                                    ('CorrelationToPower', {'ell': 2})])
def test_reference_accepts_engine_instance(cls, kw):
    """Zero-patch route 1: the reference's own FFTlog classes run on ``engine=CudaFFTEngine(...)`` (ref fftlog.py:117-132, 231, 663)."""
    ref = reference('fftlog')
    k, pk = lhs_pk(6, 1024)
    nparallel = len(kw['ell']) if isinstance(kw.get('ell'), list) else 1
    fun = pk[:, None, :] if nparallel > 1 else pk
    eng = F.CudaFFTEngine(2048, nparallel=nparallel)
    s, xi = getattr(ref, cls)(k, engine=eng, **kw)(fun)
    s2, xi2 = getattr(ref, cls)(k, engine='numpy', **kw)(fun)
    assert xi.shape == xi2.shape and np.array_equal(s, s2)
    # unfused route (rfft and irfft as two library calls around the reference's own numpy arithmetic): the scale-aware metric of SURVEY 8d
    obj = getattr(ref, cls)(k, engine='numpy', **kw)
    post = np.abs(obj.padded_postfactor[..., obj.padded_size_out_left:obj.padded_size_out_left + obj.size])
    post = post if nparallel > 1 else post[0]
    err = np.max(np.abs(xi - xi2) / post, axis=-1) / np.max(np.abs(xi2) / post, axis=-1)
    assert np.max(err) < 1e-12, err


def test_reference_fftlog_vs_cuda_engine_same_call():
    """The same user call against both packages: reference(engine='numpy') vs cosmoprimo_b200(engine='cuda'), gate of north_star (1e-10 scale-aware)."""
    ref = reference('fftlog')
    k, pk = lhs_pk(64, 2048, seed=9)
    fun = S.kaiser_multipoles(pk, np.full(pk.shape[0], 0.76))
    s2, xi2 = ref.PowerToCorrelation(k, ell=[0, 2, 4], engine='numpy')(fun)
    obj = F.PowerToCorrelation(k, ell=[0, 2, 4], engine='cuda')
    s, xi = obj(fun)
    assert np.array_equal(s, s2)
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + obj.size]
    w = 1. / np.abs(post)
    err = np.max(np.abs(xi - xi2) * w, axis=-1) / np.max(np.abs(xi2) * w, axis=-1)
    assert np.max(err) < 1e-10, np.max(err)


def reference_interpolator(ncols, seed, ntab=512):
    ref = reference('interpolator')
    ktab = np.geomspace(1e-5, 1e2, ntab)
    pk = S.eh_pk(ktab, S.lhs_cosmologies(ncols, seed=seed)).T
    return ref.PowerSpectrumInterpolator1D(ktab, pk)


def box_report(boxes, boxes_ref, dd=None):
    bad = np.nonzero(np.any(np.asarray(boxes) != np.asarray(boxes_ref), axis=1))[0]
    return bad, 'box mismatches in {} of {} columns: {}'.format(bad.size, len(boxes_ref), [(int(c), np.asarray(boxes)[c].tolist(), np.asarray(boxes_ref)[c].tolist()) for c in bad[:8]])


def reference_boxes(filt):
    """The four box indices the reference chose, recovered from its debug attributes: the reference keeps the second derivatives
    (`_dd_even`, `_dd_odd`, ref bao_filter.py:385-386) and applies ref:387-399 to them; redo exactly that on its own arrays."""
    margin_first, margin_second, offset = 20, 5, (-10, 20)
    out = []
    for dd in (filt._dd_even, filt._dd_odd):
        first = np.argmax(dd[margin_first:-margin_first], axis=0) + margin_first
        b = np.empty((dd.shape[1], 2), dtype='i4')
        for c in range(dd.shape[1]):
            second = first[c] + margin_second + np.argmax(dd[first[c] + margin_second:-margin_first, c])
            b[c] = first[c] + offset[0], second + offset[1]
        out.append(b)
    return np.concatenate(out, axis=1)


def test_register_in_reference():
    """Zero-patch route 2: after register_in_reference() the reference's own factory builds our filter for engine='wallish2018_cuda' and it
    agrees with the reference's engine='wallish2018' on the identical interpolator (ref bao_filter.py:22-31, 361-423, 912-921)."""
    reference()                                # puts baseline/_ref on the path
    refb = B.register_in_reference()
    interp = reference_interpolator(33, seed=21)
    ours = refb.PowerSpectrumBAOFilter(interp, engine='wallish2018_cuda')
    theirs = refb.PowerSpectrumBAOFilter(interp, engine='wallish2018')
    assert type(ours) is B.Wallish2018PowerSpectrumBAOFilter and type(theirs).__module__ == 'cosmoprimo.bao_filter'
    assert np.array_equal(ours.k, theirs.k) and np.array_equal(ours.pk, theirs.pk)
    assert ours.pknow.shape == theirs.pknow.shape == (1024, 33)
    bad, msg = box_report(ours._boxes, reference_boxes(theirs))
    assert bad.size == 0, msg
    assert np.max(np.abs(ours.pknow / theirs.pknow - 1.)) < 1e-10
    # the derived objects the reference builds from a filter work on ours as well (ref:115-145)
    smooth = ours.smooth_pk_interpolator()
    assert type(smooth).__module__ == 'cosmoprimo.interpolator'
    np.testing.assert_allclose(smooth(ours.k[10:-10]), ours.pknow[10:-10], rtol=1e-9)
    with pytest.raises(ValueError):
        refb.PowerSpectrumBAOFilter(interp, engine='wallish2018_nonexistent')


def test_wallish_parity_at_scale_vs_reference():
    """over 4096 Latin-hypercube spectra run through the reference itself, COUNT the columns whose boxes differ (expected 0, any
    mismatch is listed) and require max |pknow/ref - 1| <= 1e-10 on all the others."""
    refb = reference('bao_filter')
    ncols = 4096
    interp = reference_interpolator(ncols, seed=42)
    theirs = refb.PowerSpectrumBAOFilter(interp, engine='wallish2018')
    ours = B.PowerSpectrumBAOFilter(interp, engine='wallish2018_cuda')
    bad, msg = box_report(ours._boxes, reference_boxes(theirs))
    print('wallish_box_mismatches = {} of {}'.format(bad.size, ncols))
    assert bad.size == 0, msg
    err = np.max(np.abs(ours.pknow / theirs.pknow - 1.), axis=0)
    assert np.max(err) < 1e-10, (np.max(err), int(np.argmax(err)))
