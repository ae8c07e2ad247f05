"""Stress: host-path results (pageable numpy input, forced persistent kernels on the chunks) against the device path, many repetitions with
other host-path calls in between; prints the number of calls with wrong rows."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation

cases = []
for n, B, ells, kern in [(4096, 700, [1], 'pp'), (2048, 601, [0, 2, 4], 'stream'), (1024, 2001, [0, 2], 'pp'), (4096, 300, [0, 2], 'pp')]:
    k = np.geomspace(1e-5, 1e2, n)
    pk = S.eh_pk(k, S.lhs_cosmologies(B, seed=n + B))
    fun = np.repeat(pk[:, None, :], len(ells), axis=1) * (1. + np.arange(len(ells)))[None, :, None]
    obj = PowerToCorrelation(k, ell=ells)
    os.environ['CPF_FFTLOG_KERNEL'] = 'fast'
    ref = obj(torch.from_numpy(fun).cuda())[1].cpu().numpy()
    cases.append((n, B, kern, obj, fun, ref))
nbad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    for n, B, kern, obj, fun, ref in cases:
        os.environ['CPF_FFTLOG_KERNEL'] = kern
        out = obj(np.array(fun))[1]
        bad = np.nonzero(np.any(np.abs(out - ref) > 1e-9 * np.max(np.abs(ref), axis=-1, keepdims=True), axis=(1, 2)))[0]
        if bad.size:
            nbad += 1
            print('rep', rep, 'n', n, 'B', B, kern, 'bad rows', bad[:24], bad.size, flush=True)
print('calls with wrong rows:', nbad)
