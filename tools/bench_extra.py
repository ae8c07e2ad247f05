"""Secondary measurements on one GPU (the other BASELINE.json configs); the headline line is bench.py's.
Prints one JSON object.  `--quick`: only a small Wallish2018 run (used under ncu)."""
import os
import sys
import json
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from cosmoprimo_b200 import synthetic as S, _lib
from cosmoprimo_b200.fftlog import PowerToCorrelation, CorrelationToPower, TophatVariance
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D
from cosmoprimo_b200.bao_filter import PowerSpectrumBAOFilter
from cosmoprimo_b200.eisenstein_hu import EisensteinHu
from cosmoprimo_b200.interp import spline_eval_rows


def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not os.environ.get('CPF_BENCH_NO_PREQUEUE'):
        torch.cuda._sleep(800000)     # the first launch is queued before the start event fires (bench.py::hold_stream)
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def wallish(ncols, reps):
    ktab = np.geomspace(1e-5, 1e2, 512)
    base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
    pk = torch.from_numpy(np.tile(base, (1, ncols // 256)) * (1 + 1e-3 * np.arange(ncols) / ncols)).cuda()
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
    t = timed(lambda: filt(interp), reps=reps, warm=1)
    # kernel-only: the C entry point on pre-evaluated arrays
    klin = np.linspace(interp.extrap_kmin, 2., 4096)
    pklin, pkout = interp(klin), interp(filt.k)
    lib = _lib.load()
    kl, ko = torch.from_numpy(klin).cuda(), torch.from_numpy(filt.k).cuda()
    out = torch.empty_like(pkout)
    stream = torch.cuda.current_stream().cuda_stream
    tk = timed(lambda: _lib.check(lib.cpf_wallish2018(kl.data_ptr(), pklin.data_ptr(), 4096, ko.data_ptr(), pkout.data_ptr(), filt.k.size, ncols,
                                                      out.data_ptr(), None, 1, 0, stream)), reps=reps, warm=1)
    return {'ncols': ncols, 'pk_per_s_with_spline_eval': ncols / t, 'pk_per_s_filter_only': ncols / tk, 'algorithmic_gb_s_filter_only': ncols * 40960 / tk / 1e9}


def main():
    quick = '--quick' in sys.argv
    res = {}
    if quick:
        res['wallish2018'] = wallish(2048, 2)
        print(json.dumps(res))
        return
    for n, B in [(2048, 100000), (1024, 100000)]:
        k = np.geomspace(1e-5, 1e2, n)
        base = S.eh_pk(k, S.lhs_cosmologies(1000, seed=42))
        D2 = S.growth_factor(np.linspace(0., 3., B // 1000), 0.31)**2
        fun = torch.from_numpy((base[:, None, :] * D2[None, :, None]).reshape(B, n)).cuda()
        tv = TophatVariance(k)
        t = timed(lambda: tv(fun))
        res['tophat_variance_nk%d' % n] = {'rows': B, 'transforms_per_s': B / t}
        p2x = PowerToCorrelation(k)
        s, xi = p2x(fun)
        x2p = CorrelationToPower(s)
        t = timed(lambda: x2p(p2x(fun)[1]))
        res['roundtrip_nk%d' % n] = {'rows': B, 'transforms_per_s': 2 * B / t}
        # sigma(r,z): FFTLog + spline fit + evaluation at 10 radii (BASELINE config 3, reduced batch)
        if n == 2048:
            interp = PowerSpectrumInterpolator1D(k, fun[:20000].T.contiguous(), extrap_kmin=1e-5 * (1 - 1e-9), extrap_kmax=1e2 * (1 + 1e-9))
            r = np.linspace(1., 20., 10)
            t = timed(lambda: interp.sigma_r(r, nk=2048), reps=5, warm=1)
            res['sigma_rz_nk2048'] = {'rows': 20000, 'rows_per_s': 20000 / t}
    # on-device Eisenstein-Hu generator: rows/s alone, and BASELINE config 3 at full size end to end on the
    # device: 10 000 cosmologies x 100 redshifts = 1 M rows, nk = 2048: generator -> TophatVariance -> sigma at 10 radii
    n, ncosmo, nz = 2048, 10000, 100
    k = np.geomspace(1e-5, 1e2, n)
    par = S.lhs_cosmologies(ncosmo, seed=42)
    eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    zgrid = np.linspace(0., 3., nz)[None, :]
    t = timed(lambda: eh.pk(k, z=zgrid), reps=3, warm=1)
    res['eh_generator_zgrid_nk2048'] = {'rows': ncosmo * nz, 'rows_per_s': ncosmo * nz / t, 'note': '100 redshifts per cosmology, one transfer function per cosmology'}
    eh1 = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    zz = np.linspace(0., 3., ncosmo)
    t = timed(lambda: eh1.pk(k, z=zz), reps=5, warm=1)
    res['eh_generator_nk2048'] = {'rows': ncosmo, 'rows_per_s': ncosmo / t, 'points_per_s': ncosmo * n / t, 'note': 'one redshift per cosmology'}
    tv = TophatVariance(k)
    r = np.linspace(1., 20., 10)
    s_grid = tv.y if tv.y.ndim == 1 else tv.y[0]

    def sigma_rows():
        var = tv(eh.pk(k, z=zgrid).reshape(ncosmo * nz, n))[1]
        return (spline_eval_rows(s_grid, var, r))**0.5
    t = timed(sigma_rows, reps=3, warm=1)
    res['config3_sigma_rz_1M_rows_on_device'] = {'rows': ncosmo * nz, 'rows_per_s': ncosmo * nz / t, 'seconds': t,
                                                 'note': 'EH generator + TophatVariance + windowed row splines at 10 radii, nothing crosses PCIe but 80 B per row'}
    fun = eh.pk(k, z=zgrid).reshape(ncosmo * nz, n)
    t = timed(lambda: spline_eval_rows(s_grid, tv(fun)[1], r), reps=3, warm=1)
    res['sigma_rz_rows_nk2048'] = {'rows': ncosmo * nz, 'rows_per_s': ncosmo * nz / t, 'note': 'TophatVariance + windowed row splines at 10 radii, device rows in'}
    del fun
    torch.cuda.empty_cache()
    # config 1: latency of one host-array call, nk = 1024
    k = np.geomspace(1e-5, 1e2, 1024)
    pk = S.eh_pk(k)
    f = PowerToCorrelation(k)
    f(pk)
    t0 = time.perf_counter()
    for _ in range(200): f(pk)
    res['single_call_latency_us_nk1024_host'] = (time.perf_counter() - t0) / 200 * 1e6
    t0 = time.perf_counter()
    for _ in range(50): PowerToCorrelation(k)(pk)
    res['construct_plus_call_us_nk1024_host'] = (time.perf_counter() - t0) / 50 * 1e6
    res['wallish2018'] = wallish(65536, 3)       # BASELINE config 4 at full size
    print(json.dumps(res))


if __name__ == '__main__':
    main()
