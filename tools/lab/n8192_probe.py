"""nk = 4096 (N = 8192) throughput probe: P(k) -> xi multipoles, device resident."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation
for nk, ncosmo in [(4096, 2048), (4096, 8192)]:
    k = np.geomspace(1e-5, 1e2, nk)
    fun = torch.from_numpy(S.kaiser_multipoles(S.eh_pk(k, S.lhs_cosmologies(ncosmo, seed=42)), np.full(ncosmo, 0.76))).cuda()
    obj = PowerToCorrelation(k, ell=[0, 2, 4])
    for _ in range(5): obj(fun)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): obj(fun)
    e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / 20
    print('nk=%d, %d transforms: %.3f ms, %.2f M transforms/s (roofline 64.3 M: %.3f)' % (nk, 3 * ncosmo, 1e3 * dt, 3 * ncosmo / dt / 1e6, 3 * ncosmo / dt / 64.3e6))
