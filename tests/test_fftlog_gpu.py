"""Parity of the CUDA FFTLog path (through the ctypes C ABI) with the reference: golden vectors produced by the
unmodified reference, the oracle on seeded inputs, and size-independent properties at BASELINE sizes."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, Golden, scale_aware_error
from cosmoprimo_b200 import fftlog as F, _lib, synthetic as S
from oracle import fftlog_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-10   # BASELINE.json north_star: max relative error <= 1e-10, scale-aware form of SURVEY.md §8(d)
CASES = list(range(len(load_golden().cases)))


def cropped_post(obj, g_ref):
    post = np.asarray(obj.padded_postfactor)
    if post.shape[-1] != g_ref.shape[-1]:
        post = post[..., obj.padded_size_out_left:obj.padded_size_out_left + obj.size]
    if post.shape[-1] != g_ref.shape[-1]:
        post = np.ones(g_ref.shape[-1])
    return post


def build(golden, idx):
    case = golden.cases[idx]
    obj = getattr(F, case['cls'])(golden.inp(case['grid']), **Golden.ctor_kwargs(case))
    if case['inv']:
        obj.inv()
    return obj


PLANNERS = {'HankelTransform': O.plan_hankel, 'PowerToCorrelation': O.plan_power_to_correlation,
            'CorrelationToPower': O.plan_correlation_to_power, 'TophatVariance': O.plan_tophat_variance,
            'GaussianVariance': O.plan_gaussian_variance}


def oracle_full_output(golden, idx):
    case = golden.cases[idx]
    pl = PLANNERS[case['cls']](golden.inp(case['grid']), **Golden.ctor_kwargs(case))
    kw = Golden.call_kwargs(case)
    kw['keep_padding'] = True
    return O.execute(pl, golden.fun(idx), **kw)[1]


@pytest.mark.parametrize('idx', CASES)
def test_golden(fftlog_golden, idx):
    case = fftlog_golden.cases[idx]
    obj = build(fftlog_golden, idx)
    y, g = obj(fftlog_golden.fun(idx), **Golden.call_kwargs(case))
    y_ref, g_ref = fftlog_golden.get(idx, 'y'), fftlog_golden.get(idx, 'g')
    assert isinstance(g, np.ndarray) and g.shape == g_ref.shape and g.dtype == g_ref.dtype, case
    np.testing.assert_allclose(y, y_ref, rtol=1e-14, atol=0)
    post = cropped_post(obj, g_ref)
    err = scale_aware_error(g, g_ref, post)
    kw = Golden.call_kwargs(case)
    if kw.get('extrap', 0) != 0 and not kw.get('keep_padding', False):
        # non-zero padding: the (biased) padded input, hence the FFT's rounding error, is orders of magnitude larger than
        # the cropped output; normalise by the scale of the whole padded output instead (oracle, keep_padding=True)
        full = oracle_full_output(fftlog_golden, idx)
        w = 1. / np.abs(np.asarray(obj.padded_postfactor))
        scale = np.max(np.abs(full) * w, axis=-1)
        err = np.max(np.max(np.abs(g - g_ref) / np.abs(post), axis=-1) / scale)
    assert err < TOL, (case, err)
    assert err < 1e-12, (case, err)   # ~1e-15 typically


def test_analytic_hankel_pair():
    """Reference KAT (tests/test_fftlog.py:56-89) through the cuda engine, incl. inv() and a batched (3, 60) input."""
    ffun = lambda x: 1 / (1 + x**2)**1.5
    gfun = lambda y: np.exp(-y)
    x = np.logspace(-3, 3, num=60, endpoint=False)
    f = ffun(x)
    hf = F.HankelTransform(x, nu=0, q=1, lowring=True, engine='cuda')
    y, g = hf(f, extrap='log')
    assert np.allclose(g, gfun(y), rtol=1e-8, atol=1e-8)
    hf.inv()
    x2, f2 = hf(g, extrap='log')
    assert np.allclose(f2, f, rtol=1e-7, atol=1e-7)
    y = np.logspace(-4, 2, num=60, endpoint=False)
    hg = F.HankelTransform(y, nu=0, q=1, lowring=True)
    x, f = hg(gfun(y), extrap='log')
    assert np.allclose(f, ffun(x), rtol=1e-10, atol=1e-10)
    yy = np.array([y] * 3)
    scales = np.linspace(1., 3., 3)
    x, f = hg(gfun(yy) * scales[:, None], extrap='log')
    assert x.shape == (60, ) and f.shape == (3, 60)
    assert np.allclose(f / scales[:, None], ffun(x), rtol=1e-10, atol=1e-10)


def test_power_to_correlation_roundtrip(fftlog_golden):
    """Mirror of the reference's test_power_to_correlation (tests/test_fftlog.py:92-109)."""
    k, pk = fftlog_golden.inp('k1000'), fftlog_golden.inp('pk1000')
    multipoles = []
    ells = [0, 1, 2, 3, 4]
    for ell in ells:
        s, xi = F.PowerToCorrelation(k, ell=ell, lowring=True, complex=False)(pk)
        assert xi.shape == (1000, )
        k2, pk2 = F.CorrelationToPower(s, ell=ell, lowring=True, complex=False)(xi)
        idx = (1e-2 < k2) & (k2 < 10.)
        assert np.allclose(pk2[idx], np.interp(k2[idx], k, pk), rtol=1e-2)
        multipoles.append(xi)
    assert np.allclose(F.PowerToCorrelation(k, ell=ells, lowring=True, q=0, complex=False)(pk)[-1], multipoles, rtol=1e-9, atol=0)
    s, xi = F.PowerToCorrelation(k, ell=0, lowring=False)(pk)
    assert np.allclose(s[::-1] * k, 1.)


def lhs_pk(B, n):
    k = np.geomspace(1e-5, 1e2, n)
    return k, S.eh_pk(k, S.lhs_cosmologies(B, seed=42))


@pytest.mark.parametrize('n,B', [(2048, 257), (1024, 64), (1000, 33), (512, 7), (300, 5), (4096, 9), (100, 3)])
def test_oracle_seeded_batch(n, B):
    """Odd batches of Latin-hypercube EH spectra vs the oracle on the same inputs; covers the three fast-path radices
    (N = 1024, 2048, 4096), windows narrower than N/2 (n = 1000, 300) and the generic kernel (N = 8192, 256)."""
    k, pk = lhs_pk(B, n)
    for cls, planner, kw in [(F.PowerToCorrelation, O.plan_power_to_correlation, dict(ell=2)),
                             (F.TophatVariance, O.plan_tophat_variance, dict())]:
        obj = cls(k, **kw)
        y, g = obj(pk)
        pl = planner(k, **kw)
        y_ref, g_ref = O.execute(pl, pk)
        assert g.shape == (B, n)
        assert np.allclose(y, y_ref, rtol=1e-14, atol=0)
        assert scale_aware_error(g, g_ref, cropped_post(obj, g_ref)) < 1e-13


def test_multipoles_config2_shapes():
    """BASELINE config 2 at reduced batch: (B,3,n) per-ell inputs and the broadcast forms (B,1,n), (n,)."""
    B, n = 17, 2048
    k, pk = lhs_pk(B, n)
    multi = S.kaiser_multipoles(pk, np.full(B, 0.76))
    obj = F.PowerToCorrelation(k, ell=[0, 2, 4])
    pl = O.plan_power_to_correlation(k, ell=[0, 2, 4])
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    for fun in [multi, pk[:, None, :], pk[0], multi[0], pk[:1][:, None, :], pk.reshape(1, B, 1, n)]:
        s, xi = obj(fun)
        s_ref, xi_ref = O.execute(pl, fun)
        assert s.shape == (3, n) and xi.shape == xi_ref.shape
        assert scale_aware_error(xi, xi_ref, post) < 1e-13
    # multi-ell == per-ell (tests/test_fftlog.py:107)
    xi = obj(multi)[1]
    for i, ell in enumerate([0, 2, 4]):
        one = F.PowerToCorrelation(k, ell=ell)(multi[:, i])[1]
        assert scale_aware_error(one, xi[:, i], post[i]) < 1e-14


@pytest.mark.parametrize('kernel,n,B,ells,per_ell', [
    ('stream', 2048, 601, [0, 2, 4], True),     # full window, odd batch, one input row per ell
    ('stream', 2048, 2001, [0, 2, 4], True),    # enough pairs per plan row for the dynamically scheduled TMA variant, odd batch
    ('stream', 2048, 2600, [1], False),         # dynamic scheduling, one plan row shared by all CTAs
    ('stream', 2048, 300, [0, 2], False),       # the same row for every ell (no P axis in the input)
    ('stream', 2000, 77, [0, 2, 4], True),      # window narrower than N/2: masked loads and stores
    ('stream', 1919, 40, [1], False),           # odd n, single plan row
    ('stream', 2048, 2, [0, 2, 4], True),       # as many CTAs as plan rows
    ('stream', 2048, 1, [0, 1, 2, 3, 4], True),  # fewer CTAs than plan rows: a CTA walks over several plan rows
    ('pp', 1024, 333, [0, 2], True),            # N = 2048: ping-pong kernel, four groups per CTA
    ('pp', 1024, 12001, [0, 2], True),          # enough pairs for the dynamically scheduled TMA variant, odd batch
    ('pp', 512, 19000, [1], False),             # N = 1024, dynamic scheduling, one plan row
    ('pp', 1000, 65, [2], False),
    ('pp', 512, 129, [0, 2, 4], True),          # N = 1024: eight groups per CTA
    ('pp', 2048, 67, [0, 2], True),             # N = 4096 through the ping-pong kernel
    ('pp', 4096, 35, [0, 2, 4], True),          # N = 8192: persistent two-chain kernel (fftlog_pp8k_kernel), TMA-staged rows, odd batch
    ('pp', 4096, 700, [1], False),              # ... several pairs per CTA, one plan row, the same row for every ell
    ('pp', 3000, 11, [0, 2], True),             # ... window narrower than N/2: direct masked loads and stores
    ('auto', 4096, 1301, [0, 2], True),         # ... chosen automatically for a large launch
])
def test_persistent_kernels(monkeypatch, kernel, n, B, ells, per_ell):
    """The persistent kernels (stream for N = 4096, ping-pong for N = 2048 / 1024) are chosen automatically for large
    launches only; here they are forced on small odd batches and compared with the oracle and with the per-pair kernel."""
    k, pk = lhs_pk(B, n)
    fun = S.kaiser_multipoles(pk, np.full(B, 0.76))[:, :len(ells)] if per_ell else pk
    if per_ell and len(ells) > 3:
        fun = np.concatenate([fun, fun[:, :len(ells) - 3] * 0.5], axis=1)
    obj = F.PowerToCorrelation(k, ell=ells)
    monkeypatch.setenv('CPF_FFTLOG_KERNEL', 'fast')
    ref_fast = obj(fun if per_ell else fun[:, None, :])[1]
    monkeypatch.setenv('CPF_FFTLOG_KERNEL', kernel)
    if kernel == 'pp' and B > 10000:
        monkeypatch.setenv('CPF_PP_DYNAMIC', '1')       # the opt-in ticket-counter variant of the ping-pong kernel
    arg = fun if per_ell else fun[:, None, :]
    if kernel == 'auto':                                # device-resident rows: one launch, large enough for the automatic choice
        import torch
        s, xi = obj(torch.from_numpy(np.ascontiguousarray(arg)).cuda())
        xi = xi.cpu().numpy()
    else:
        s, xi = obj(arg)
    assert xi.shape == (B, len(ells), n) and np.isfinite(xi).all()
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    bad = np.nonzero(np.any(np.abs(xi - ref_fast) > 1e-9 * np.max(np.abs(ref_fast), axis=-1, keepdims=True), axis=(1, 2)))[0]
    assert scale_aware_error(xi, ref_fast, post) < 1e-13, 'rows that differ from the per-pair kernel: {} ({} of {})'.format(bad[:32], bad.size, B)
    rows = sorted(set([0, B // 2, B - 1]))
    ref = O.execute(O.plan_power_to_correlation(k, ell=ells), (fun if per_ell else fun[:, None, :])[rows])[1]
    assert scale_aware_error(xi[rows], ref, post) < 1e-13


@pytest.mark.parametrize('kernel,n,B,ells', [
    ('stream', 2048, 37, [0, 2, 4]),      # persistent stream kernel, odd batch (the last pair has one row)
    ('pp', 1024, 20, [0, 2]),             # ping-pong kernel
    ('fast', 2048, 9, [0]),               # per-pair kernel
    ('fast', 1000, 6, [0, 2]),
    ('auto', 60, 7, [0]),                 # generic shared-memory kernel (N = 128)
    ('auto', 4096, 7, [0, 2]),            # N = 8192: split kernel (two 4096-point FFTs per transform)
    ('pp', 4096, 7, [0, 2]),              # N = 8192: persistent two-chain kernel
])
def test_non_finite_rows_stay_in_their_row(monkeypatch, kernel, n, B, ells):
    """The reference transforms rows independently (numpy.fft along the last axis): a NaN / Inf sample turns ITS row into NaN
    and leaves every other row untouched.  The kernels pack two rows into one complex FFT, so this has to be enforced."""
    torch = pytest.importorskip('torch')
    k, pk = lhs_pk(B, n)
    fun = np.repeat(pk[:, None, :], len(ells), axis=1) * (1. + np.arange(len(ells)))[None, :, None]
    obj = F.PowerToCorrelation(k, ell=ells)
    monkeypatch.setenv('CPF_FFTLOG_KERNEL', kernel)
    clean = obj(fun)[1]
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    dirty = fun.copy()
    bad = [(0, 0, 0, np.nan), (3, len(ells) - 1, n // 2, np.inf), (B - 1, 0, n - 1, -np.inf)]       # rows a, b of pairs, odd tail
    if B > 5:
        bad += [(4, 0, 5, np.nan), (5, 0, 7, np.nan)]                                                # both rows of one pair
    for b, p, i, val in bad:
        dirty[b, p, i] = val
    for arr in (dirty, torch.from_numpy(dirty).cuda()):
        out = obj(arr)[1]
        out = out.cpu().numpy() if isinstance(out, torch.Tensor) else out
        hit = np.zeros((B, len(ells)), dtype='?')
        for b, p, i, val in bad:
            hit[b, p] = True
        assert np.isnan(out[hit]).all()
        # the partner of a poisoned row is now transformed next to zeros instead of next to its neighbour: same numbers up
        # to the rounding of the complex arithmetic
        assert np.isfinite(out[~hit]).all() and scale_aware_error(out[~hit], clean[~hit], post[np.nonzero(~hit)[1]]) < 1e-13
    # with extrapolated padding the pads of a poisoned row are poisoned too, and stay in that row
    if kernel in ('fast', 'auto'):
        clean = obj(fun, extrap='log')[1]
        out = obj(dirty, extrap='log')[1]
        assert np.isnan(out[hit]).all() and scale_aware_error(out[~hit], clean[~hit], post[np.nonzero(~hit)[1]]) < 1e-13


def test_kernel_family_selection():
    lib = _lib.load()
    k = np.geomspace(1e-5, 1e2, 2048)
    fam = lambda obj, *a: lib.cpf_plan_kernel_family(obj._device_plan(0).handle, *a)
    obj = F.PowerToCorrelation(k)
    assert fam(obj, 0, 0., 0, 0., 0) == 2              # zero padding, cropped: pruned register kernel
    assert fam(obj, _lib.EXTRAP_LOG, 0., 0, 0., 0) == 1  # any other option: full register kernel
    assert fam(obj, 0, 0., 0, 0., 1) == 1
    assert fam(obj, 0, 1., 0, 0., 0) == 1
    assert fam(F.PowerToCorrelation(np.geomspace(1e-5, 1e2, 1000)), 0, 0., 0, 0., 0) == 2
    assert fam(F.PowerToCorrelation(np.geomspace(1e-5, 1e2, 4096)), 0, 0., 0, 0., 0) == 2   # N=8192: two 4096-point register FFTs (split kernel), pruned
    assert fam(F.PowerToCorrelation(np.geomspace(1e-5, 1e2, 4096)), _lib.EXTRAP_LOG, 0., 0, 0., 0) == 1
    assert fam(F.PowerToCorrelation(np.geomspace(1e-5, 1e2, 3000)), 0, 0., 0, 0., 0) == 2   # n = 3000 -> N = 8192
    assert fam(F.PowerToCorrelation(np.geomspace(1e-5, 1e2, 60)), 0, 0., 0, 0., 0) == 0


def test_device_buffers_match_host_path():
    torch = pytest.importorskip('torch')
    B, n = 31, 2048
    k, pk = lhs_pk(B, n)
    obj = F.PowerToCorrelation(k, ell=[0, 2, 4])
    fun = S.kaiser_multipoles(pk, np.full(B, 0.76))
    s, xi = obj(fun)
    s2, xi2 = obj(torch.from_numpy(fun).cuda())
    assert isinstance(xi2, torch.Tensor) and xi2.is_cuda and xi2.dtype == torch.float64 and tuple(xi2.shape) == xi.shape
    assert np.array_equal(xi2.cpu().numpy(), xi)       # same kernel, same bits
    # non-contiguous / float32 device input is promoted like the reference promotes numpy input
    x32 = torch.from_numpy(fun.astype('f4')).cuda()
    xi3 = obj(x32)[1]
    xi3_ref = O.execute(O.plan_power_to_correlation(k, ell=[0, 2, 4]), fun.astype('f4'))[1]
    assert xi3.dtype == torch.float64
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    assert scale_aware_error(xi3.cpu().numpy(), xi3_ref, post) < 1e-13
    # __cuda_array_interface__ producer that is not a torch tensor

    class CAI(object):
        def __init__(self, t):
            self.t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__
    xi4 = obj(CAI(torch.from_numpy(fun).cuda()))[1]
    assert np.array_equal(xi4.cpu().numpy(), xi)
    # complex post-factor on device buffers
    objc = F.PowerToCorrelation(k, ell=[0, 1], complex=True)
    xc = objc(torch.from_numpy(pk).cuda()[:, None, :])[1]
    xc_ref = O.execute(O.plan_power_to_correlation(k, ell=[0, 1], complex=True), pk[:, None, :])[1]
    assert xc.dtype == torch.complex128
    postc = objc.padded_postfactor[:, objc.padded_size_out_left:objc.padded_size_out_left + n]
    assert scale_aware_error(xc.cpu().numpy(), xc_ref, postc) < 1e-13


@pytest.mark.parametrize('env', [{}, {'CPF_STAGE_CAP_KB': '2048', 'CPF_STAGE_SMALL_KB': '512'}, {'CPF_STAGE_CAP_KB': '3072', 'CPF_STAGE_NBUF': '1'},
                                 {'CPF_STAGE_CAP_KB': '1024', 'CPF_STAGE_NBUF': '2'}])
def test_pageable_and_pinned_host_input_give_the_same_bits(monkeypatch, env):
    """Host arrays: an ordinary (pageable) numpy array is copied into page-locked bounce buffers by worker threads one chunk ahead of its
    H2D copy; a page-locked array goes straight to the copy engine.  Same kernels on the same rows => identical bits, whatever the chunking
    (many chunks, ring of one / two / four buffers, odd batch)."""
    torch = pytest.importorskip('torch')
    for name, val in env.items():
        monkeypatch.setenv(name, val)
    B, n = 341, 2048                                     # 341 x 3 x 2048 doubles = 16.8 MB in, odd batch
    k, pk = lhs_pk(64, n)
    fun = np.tile(S.kaiser_multipoles(pk, np.full(64, 0.76)), (6, 1, 1))[:B] * (1. + 1e-3 * np.arange(B))[:, None, None]
    obj = F.PowerToCorrelation(k, ell=[0, 2, 4])
    pinned = torch.from_numpy(fun).pin_memory().numpy()
    xi_dev = obj(torch.from_numpy(fun).cuda())[1].cpu().numpy()
    xi_pageable = obj(np.array(fun))[1]
    xi_pinned = obj(pinned)[1]
    assert np.array_equal(xi_pageable, xi_pinned)
    # the device path takes the persistent kernel, the chunks of the host path the per-pair one: same arithmetic up to twiddle rounding
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    assert scale_aware_error(xi_pageable, xi_dev, post) < 1e-13
    ref = O.execute(O.plan_power_to_correlation(k, ell=[0, 2, 4]), fun[[0, 170, 340]])[1]
    assert scale_aware_error(xi_pageable[[0, 170, 340]], ref, post) < 1e-13


def test_empty_and_single():
    k, pk = lhs_pk(2, 1024)
    obj = F.PowerToCorrelation(k)
    s, xi = obj(np.empty((0, 1024)))
    assert xi.shape == (0, 1024)
    s, xi = obj(pk[:1])
    assert xi.shape == (1, 1024)
    assert scale_aware_error(xi[0], O.execute(O.plan_power_to_correlation(k), pk[0])[1], cropped_post(obj, xi[0])) < 1e-13


def test_tables_are_revalidated_after_mutation():
    """Subclasses and inv() rewrite the public tables after the engine exists (fftlog.py:117, 243-248, 319-330): the
    device plan must follow."""
    k, pk = lhs_pk(3, 1024)
    obj = F.PowerToCorrelation(k)
    xi = obj(pk)[1]
    obj.padded_prefactor *= 2.
    assert np.allclose(obj(pk)[1], 2 * xi, rtol=1e-14, atol=0)
    obj.padded_prefactor /= 2.
    s = obj.y[0].copy()
    obj.inv()
    k2, pk2 = obj(xi)
    ref = O.plan_power_to_correlation(k)
    O.invert_plan(ref)
    assert scale_aware_error(pk2, O.execute(ref, xi)[1], cropped_post(obj, pk2)) < 1e-10


def test_linearity_and_roundtrip_full_size():
    """BASELINE config 2 at full size (4096 cosmologies x ell=0,2,4, nk=2048), through size-independent properties:
    linearity of the transform, and xi -> P -> xi round trip (config 5's pattern)."""
    B, n = 4096, 2048
    k, pk = lhs_pk(B, n)
    fun = S.kaiser_multipoles(pk, np.full(B, 0.76))
    obj = F.PowerToCorrelation(k, ell=[0, 2, 4])
    s, xi = obj(fun)
    assert xi.shape == (B, 3, n) and np.isfinite(xi).all()
    post = obj.padded_postfactor[:, obj.padded_size_out_left:obj.padded_size_out_left + n]
    perm = np.random.default_rng(0).permutation(B)
    a, b = 0.37, -1.83
    lin = obj(a * fun + b * fun[perm])[1]
    assert scale_aware_error(lin, a * xi + b * xi[perm], post) < 1e-13
    # spot-check rows against the oracle
    rows = [0, 1, 2047, 4095]
    ref = O.execute(O.plan_power_to_correlation(k, ell=[0, 2, 4]), fun[rows])[1]
    assert scale_aware_error(xi[rows], ref, post) < 1e-13
    back = F.CorrelationToPower(s, ell=[0, 2, 4])
    k2, pk2 = back(xi)
    idx = (1e-2 < k2[0]) & (k2[0] < 10.)
    for i in range(3):
        interp = np.array([np.interp(k2[i][idx], k, fun[r, i]) for r in rows])
        assert np.allclose(pk2[rows, i][:, idx], interp, rtol=1e-2)


def test_unfused_engine_duck_type():
    """CudaFFTEngine.forward/backward == NumpyFFTEngine's (fftlog.py:538-544)."""
    rng = np.random.default_rng(1)
    for size in [8, 128, 2048, 4096, 8192]:
        eng = F.CudaFFTEngine(size)
        x = rng.standard_normal((5, size))
        X = eng.forward(x)
        assert X.shape == (5, size // 2 + 1) and X.dtype == np.complex128
        Xr = np.fft.rfft(x, axis=-1)
        assert np.max(np.abs(X - Xr)) < 1e-13 * np.max(np.abs(Xr))
        c = rng.standard_normal((3, 2, size // 2 + 1)) + 1j * rng.standard_normal((3, 2, size // 2 + 1))
        g = eng.backward(c)
        gr = np.fft.irfft(c.conj(), n=size, axis=-1)
        assert g.shape == (3, 2, size) and np.max(np.abs(g - gr)) < 1e-13 * np.max(np.abs(gr))


TICKET_STRESS = r'''
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from cosmoprimo_b200 import fftlog as F, synthetic as S, _lib
lib = _lib.load()
n, B = 2048, 2600
k = np.geomspace(1e-5, 1e2, n)
pk = torch.from_numpy(S.eh_pk(k, S.lhs_cosmologies(B, seed=5))).cuda()
obj = F.PowerToCorrelation(k, ell=[1])
ref = obj(pk)[1].clone()
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(8)]
outs = []
torch.cuda._sleep(int(1.5e9))              # keep the GPU busy (~0.75 s) while the launches below are queued: all of them are in flight at once
for st in streams:
    st.wait_stream(torch.cuda.current_stream())
for rep in range(6):                       # 48 launches queued on 8 streams behind the sleep kernel
    for st in streams:
        with torch.cuda.stream(st):
            outs.append(obj(pk)[1])
torch.cuda.synchronize()
bad = sum(int(not torch.equal(o, ref)) for o in outs)
print('RESULT', bad, lib.cpf_counter(0), lib.cpf_counter(1))
'''


@pytest.mark.parametrize('slots', [2, 64])
def test_ticket_ring_with_more_launches_in_flight_than_slots(slots):
    """VERDICT r1 'weak' 13: launches that would share a ticket slot with a launch still in flight must not skip pairs.  With a ring of 2 slots
    and 48 launches queued on 8 streams most launches find their slot busy and take the static split; all results are bit-identical."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CPF_TICKET_SLOTS=str(slots), CPF_FFTLOG_KERNEL='stream')
    res = subprocess.run([sys.executable, '-c', TICKET_STRESS.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    bad, dynamic, fallback = (int(x) for x in res.stdout.strip().splitlines()[-1].split()[1:])
    assert bad == 0 and dynamic >= 1 and dynamic + fallback == 49
    if slots == 2:
        assert fallback > 0        # the guard was exercised


def test_scratch_pool_is_private_and_trimmable():
    """the library must not touch the attributes of the device's default memory pool; its own pool is bounded and cpf_trim empties it."""
    torch = pytest.importorskip('torch')
    lib = _lib.load()
    k, pk = lhs_pk(512, 2048)
    free0 = torch.cuda.mem_get_info()[0]
    F.PowerToCorrelation(k)(pk)                     # host path: staging buffers come from the private pool
    _lib.check(lib.cpf_trim(0))
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20                 # plan tables and the context only: the staging scratch went back to the driver
    assert lib.cpf_counter(99) == -1


def test_inv_of_complex_transform_raises_like_the_reference():
    """inv() of a complex=True plan makes the pre-factor complex; numpy.fft.rfft refuses the reference's complex product (TypeError, ref fftlog.py:540)."""
    k = np.geomspace(1e-5, 1e2, 256)
    obj = F.PowerToCorrelation(k, ell=1, complex=True)
    s, xi = obj(1. / k)
    assert xi.dtype == np.complex128
    obj.inv()
    with pytest.raises(TypeError):
        obj(xi.real)


@pytest.mark.parametrize('n,B,kw,callkw', [
    (4096, 33, {'ell': [0, 2, 4]}, {}),                                        # N = 8192, cropped, zero padding: pruned variant, odd batch
    (4096, 4, {'ell': 1}, {'extrap': 'log'}),                                  # extrapolated padding: all 8192 samples are live
    (4096, 3, {'ell': [0, 2]}, {'keep_padding': True}),
    (4096, 5, {'ell': [1, 3], 'complex': True}, {}),                           # complex post-factor
    (3000, 6, {'ell': 0, 'q': 0.5}, {'extrap': ('edge', 'log')}),              # n = 3000 -> N = 8192, window not a power of two
    (3000, 2, {'ell': 2}, {}),
])
def test_n8192_split_kernel_vs_oracle(n, B, kw, callkw):
    """VERDICT r1 missing 2: nk = 4096 (ref fftlog.py:149-150 => N = 8192) runs on the register FFT (fftlog_split2_kernel), not on the radix-2
    shared-memory kernel; every option against the oracle on seeded EH spectra."""
    k, pk = lhs_pk(B, n)
    nell = len(kw['ell']) if isinstance(kw['ell'], list) else 1
    fun = pk[:, None, :] if nell > 1 else pk
    obj = F.PowerToCorrelation(k, **kw)
    assert obj.padded_size == 8192
    s, xi = obj(fun, **callkw)
    s_ref, ref = O.execute(O.plan_power_to_correlation(k, **kw), fun, **callkw)
    assert xi.shape == ref.shape and xi.dtype == ref.dtype
    np.testing.assert_allclose(s, s_ref, rtol=1e-14, atol=0)
    post = cropped_post(obj, ref)
    if callkw.get('extrap'):
        # extrapolated padding: normalise by the scale of the whole padded output, as test_golden does
        full = O.execute(O.plan_power_to_correlation(k, **kw), fun, **dict(callkw, keep_padding=True))[1]
        scale = np.max(np.abs(full) / np.abs(np.asarray(obj.padded_postfactor)), axis=-1)
        err = np.max(np.max(np.abs(xi - ref) / np.abs(post), axis=-1) / scale)
    else:
        err = scale_aware_error(xi, ref, post)
    assert err < 1e-13, err
