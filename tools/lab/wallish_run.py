"""Profiler driver: cpf_wallish2018 over N device-resident spectra, a few repetitions (python tools/lab/wallish_run.py [ncols] [reps] [rows]):
`rows` = the linear-grid spectra handed over one row per spectrum (cpf_wallish2018_rows), checked against the column entry."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cosmoprimo_b200 import synthetic as S, _lib
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ktab = np.geomspace(1e-5, 1e2, 512)
base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
pk = torch.from_numpy(np.tile(base, (1, (ncols + 255) // 256))[:, :ncols] * (1 + 1e-3 * np.arange(ncols) / ncols)).cuda()
interp = PowerSpectrumInterpolator1D(ktab, pk)
klin = np.linspace(interp.extrap_kmin, 2., 4096)
kout = np.geomspace(interp.extrap_kmin, interp.extrap_kmax, 1024)
pklin, pkout = interp(klin), interp(kout)
lib = _lib.load()
kl, ko = torch.from_numpy(klin).cuda(), torch.from_numpy(kout).cuda()
out = torch.empty_like(pkout)
stream = torch.cuda.current_stream().cuda_stream
call = lambda: _lib.check(lib.cpf_wallish2018(kl.data_ptr(), pklin.data_ptr(), 4096, ko.data_ptr(), pkout.data_ptr(), kout.size, ncols, out.data_ptr(), None, 1, 0, stream))
if len(sys.argv) > 3 and sys.argv[3] == 'rows':
    call()
    torch.cuda.synchronize()
    ref = out.clone()
    rows = interp._interp.eval_rows(klin)                      # (ncols, 4096)
    assert tuple(rows.shape) == (ncols, 4096)
    call = lambda: _lib.check(lib.cpf_wallish2018_rows(klin.ctypes.data, rows.data_ptr(), 4096, kout.ctypes.data, pkout.data_ptr(), kout.size, ncols, out.data_ptr(), None, 1, 0, stream))
    out.zero_()
    call()
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref)) or bool(((out == ref) | (torch.isnan(out) & torch.isnan(ref))).all())
    print('rows entry vs column entry: bit-identical = {}, max |diff/ref| = {:.2e}'.format(same, float(((out - ref).abs() / ref.abs()).nan_to_num(0.).max())))
call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    call()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print('wallish2018: {} spectra, {:.3f} ms per call, {:.2f} M P(k)/s'.format(ncols, dt * 1e3, ncols / dt / 1e6))
