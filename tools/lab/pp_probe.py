"""Persistent ping-pong kernels on device-resident rows: nk = 512 / 1024 (their default path) and nk = 2048 (CPF_FFTLOG_KERNEL=pp)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation

for n, B, kern in [(512, 16384, 'auto'), (1024, 8192, 'auto'), (2048, 4096, 'pp'), (2048, 4096, 'auto')]:
    os.environ['CPF_FFTLOG_KERNEL'] = kern
    k = np.geomspace(1e-5, 1e2, n)
    pk = S.eh_pk(k, S.lhs_cosmologies(B, seed=1))
    fun = torch.from_numpy(S.kaiser_multipoles(pk, np.full(B, 0.76))).cuda()
    obj = PowerToCorrelation(k, ell=[0, 2, 4], engine='cuda', device=0)
    keep = [obj(fun)[1] for _ in range(5)]
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50):
            keep[i % 5] = obj(fun)[1]
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    print('nk = {:4d} ({:6s}): {:.4f} ms per {} transforms = {:.2f} M transforms/s'.format(n, kern, best, 3 * B, 3 * B / best / 1e3))
    del fun, keep
