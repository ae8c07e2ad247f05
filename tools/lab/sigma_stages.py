"""Stage timings of the sigma(r, z) row path on one GPU (CUDA events): TophatVariance, row splines, generator."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import TophatVariance
from cosmoprimo_b200.interp import spline_eval_rows
from cosmoprimo_b200.eisenstein_hu import EisensteinHu


def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


n, rows = 2048, 100000
k = np.geomspace(1e-5, 1e2, n)
par = S.lhs_cosmologies(1000, seed=42)
eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
fun = eh.pk(k, z=np.linspace(0., 3., 100)[None, :]).reshape(rows, n)
tv = TophatVariance(k)
s_grid = tv.y if tv.y.ndim == 1 else tv.y[0]
r = np.linspace(1., 20., 10)
var = tv(fun)[1]
res = {}
res['tophat_ms'] = 1e3 * timed(lambda: tv(fun))
for w in [0, 32, 64, 128]:
    res['rows_eval_window%d_ms' % w] = 1e3 * timed(lambda: spline_eval_rows(s_grid, var, r, window=w))
res['rows_eval_1query_ms'] = 1e3 * timed(lambda: spline_eval_rows(s_grid, var, r[:1]))
res['rows_eval_1000rows_ms'] = 1e3 * timed(lambda: spline_eval_rows(s_grid, var[:1000], r))
res['both_ms'] = 1e3 * timed(lambda: spline_eval_rows(s_grid, tv(fun)[1], r))
res['generator_zgrid_ms'] = 1e3 * timed(lambda: eh.pk(k, z=np.linspace(0., 3., 100)[None, :]))
print(json.dumps(res))
