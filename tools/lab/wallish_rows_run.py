"""Profiler driver: cpf_wallish2018_rows (one row per spectrum, host grids) over N device-resident spectra (python tools/lab/wallish_rows_run.py [ncols] [reps])."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cosmoprimo_b200 import synthetic as S, _lib
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ktab = np.geomspace(1e-5, 1e2, 512)
base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
pk = torch.from_numpy(np.tile(base, (1, (ncols + 255) // 256))[:, :ncols] * (1 + 1e-3 * np.arange(ncols) / ncols)).cuda()
interp = PowerSpectrumInterpolator1D(ktab, pk)
klin = np.linspace(interp.extrap_kmin, 2., 4096)
kout = np.geomspace(interp.extrap_kmin, interp.extrap_kmax, 1024)
rows, pkout = interp._interp.eval_rows(klin), interp(kout)
lib = _lib.load()
out = torch.empty_like(pkout)
stream = torch.cuda.current_stream().cuda_stream
call = lambda: _lib.check(lib.cpf_wallish2018_rows(klin.ctypes.data, rows.data_ptr(), 4096, kout.ctypes.data, pkout.data_ptr(), kout.size, ncols, out.data_ptr(), None, 1, 0, stream))
call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    call()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print('wallish2018 (rows): {} spectra, {:.3f} ms per call, {:.2f} M P(k)/s'.format(ncols, dt * 1e3, ncols / dt / 1e6))
