"""Timing of the log-log spline construction and evaluation on the bench's spline_fit_and_eval workload (540 knots x N spectra -> 2048 wavenumbers),
construction and evaluation timed separately (r6c: tile-kernel A/B, r6d: fast_log10 / fast_exp10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.interp import Interpolator1D
from cosmoprimo_b200.interpolator import _pad_log_knots
from oracle import spline_oracle as SO

def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps

ktab = np.geomspace(1e-4, 50., 540)
tab = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=3)).T
kq = np.geomspace(1e-4, 50., 2048)
logk, lo, hi = _pad_log_knots(ktab)
kpad = 10**np.concatenate([lo, logk, hi])
ref = SO.interpolator1d(ktab, tab[:, :32], interp_x='log', interp_fun='log', assume_sorted=True)(kq)
for nsp in [4096, 16384, 65536]:
    tab_d = torch.from_numpy(np.tile(tab, (1, nsp // 256))).cuda()
    for rnd in range(2):
        tc = t(lambda: Interpolator1D.padlog(kpad, tab_d))
        sp = Interpolator1D.padlog(kpad, tab_d)
        te = t(lambda: sp.eval_rows(kq))
        tt = t(lambda: Interpolator1D.padlog(kpad, tab_d).eval_rows(kq))
        out = sp.eval_rows(kq)[:32].cpu().numpy()
        err = float(np.max(np.abs(out / ref.T - 1.)))
        print('nsp %6d: construct %.3f ms, evaluate %.3f ms, both %.3f ms = %.1f M spectra/s = %.3f of HBM (6.55 TB/s); max rel error vs oracle %.1e'
              % (nsp, tc * 1e3, te * 1e3, tt * 1e3, nsp / tt / 1e6, nsp * 8. * (540 + 2048) / tt / 6.5558e12, err), flush=True)
    del tab_d, sp
    torch.cuda.empty_cache()
