"""The callers either side of the hot path on device tables, for a launch list (ncu --metrics gpu__time_duration.sum): interpolator construction
(log-log table with continuation knots), sigma_r (spline evaluation in rows -> TophatVariance -> row splines + sqrt) and the Wallish2018 filter with
its two input evaluations.  Every launch between the two markers should be a cpf:: kernel (no at:: elementwise / cat / index kernels)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D
from cosmoprimo_b200.bao_filter import PowerSpectrumBAOFilter

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ktab = np.geomspace(1e-4, 50., 540)
base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
pk = torch.from_numpy(np.tile(base, (1, ncols // 256))).cuda()
r = np.linspace(2., 20., 10)
for rep in range(3):
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('glue_path')
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    sig = interp.sigma_r(r)
    filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print('sigma_r', tuple(sig.shape), float(sig[1, 0]), 'pknow', tuple(filt.pknow.shape), bool(torch.isfinite(filt.pknow).all()))
