"""Timing probe for the scratch-pool settings: spline fit + eval and Wallish2018 under different CPF_SCRATCH_KEEP_MB (subprocess per setting)."""
import os
import subprocess
import sys

CODE = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
from cosmoprimo_b200 import synthetic as S, _lib
from cosmoprimo_b200.interp import Interpolator1D
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D
n = 2048
ktab = np.geomspace(1e-4, 50., 540)
tab = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=3)).T
tab_d = torch.from_numpy(np.tile(tab, (1, 16))).cuda()
kq = np.geomspace(1e-4, 50., n)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print('spline fit+eval 4096 spectra: %%.3f ms' %% (1e3 * t(lambda: Interpolator1D(ktab, tab_d, interp_x='log', interp_fun='log', assume_sorted=True).eval_rows(kq))))
for ncols in [4096, 65536]:
    ktab2 = np.geomspace(1e-5, 1e2, 512)
    base = S.eh_pk(ktab2, S.lhs_cosmologies(256, seed=42)).T
    pk = torch.from_numpy(np.tile(base, (1, ncols // 256))).cuda()
    interp = PowerSpectrumInterpolator1D(ktab2, pk)
    klin = np.linspace(interp.extrap_kmin, 2., 4096)
    kout = np.geomspace(interp.extrap_kmin, interp.extrap_kmax, 1024)
    pklin, pkout = interp(klin), interp(kout)
    lib = _lib.load()
    kl, ko = torch.from_numpy(klin).cuda(), torch.from_numpy(kout).cuda()
    out = torch.empty_like(pkout)
    st = torch.cuda.current_stream().cuda_stream
    dt = t(lambda: _lib.check(lib.cpf_wallish2018(kl.data_ptr(), pklin.data_ptr(), 4096, ko.data_ptr(), pkout.data_ptr(), kout.size, ncols, out.data_ptr(), None, 1, 0, st)), reps=3)
    print('wallish %%d spectra: %%.3f ms = %%.2f M/s' %% (ncols, 1e3 * dt, ncols / dt / 1e6))
    dt = t(lambda: (interp(klin), interp(kout)), reps=3)
    print('  two interpolator evaluations: %%.3f ms' %% (1e3 * dt))
    del pk, interp, pklin, pkout, out
    torch.cuda.empty_cache()
'''
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for keep in ['2048', '100000']:
    print('=== CPF_SCRATCH_KEEP_MB =', keep, flush=True)
    res = subprocess.run([sys.executable, '-c', CODE % root], env=dict(os.environ, CPF_SCRATCH_KEEP_MB=keep), capture_output=True, text=True)
    print(res.stdout, res.stderr[-1500:], flush=True)
