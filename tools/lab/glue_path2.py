"""Launch list of the other callers on device tables: to_xi, to_pk, the 2-D interpolator's sigma_rz (what torch-eager kernels are left?)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D, PowerSpectrumInterpolator2D

ncols = 4096
ktab = np.geomspace(1e-4, 50., 540)
base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
pk = torch.from_numpy(np.tile(base, (1, ncols // 256))).cuda()
z = np.linspace(0., 2., 16)
pk2 = torch.from_numpy(base[:, :1] * (1. + z)[None, :]**-2).cuda()
for rep in range(2):
    torch.cuda.synchronize()
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    xi = interp.to_xi()
    back = xi.to_pk()
    i2 = PowerSpectrumInterpolator2D(ktab, z, pk2)
    sig = i2.sigma_rz(np.array([4., 8.]), z)
    torch.cuda.synchronize()
print('done', tuple(sig.shape))
