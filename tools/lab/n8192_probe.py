"""nk = 4096 (N = 8192) throughput on device-resident rows: persistent two-chain kernel (auto / pp) against the split kernel (fast)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation

B, n = 2048, 4096
k = np.geomspace(1e-5, 1e2, n)
pk = S.eh_pk(k, S.lhs_cosmologies(B, seed=1))
fun = torch.from_numpy(S.kaiser_multipoles(pk, np.full(B, 0.76))).cuda()
for kern in ['fast', 'pp', 'auto']:
    os.environ['CPF_FFTLOG_KERNEL'] = kern
    obj = PowerToCorrelation(k, ell=[0, 2, 4], engine='cuda', device=0)
    out = obj(fun)[1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    for _ in range(5):
        keep.append(obj(fun)[1])
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        keep[i % 5] = obj(fun)[1]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if kern == 'fast':
        ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    print('{:5s}: {:.3f} ms per {} transforms = {:.2f} M transforms/s; max |diff| vs split kernel / max |xi| = {:.2e}'.format(kern, ms, 3 * B, 3 * B / ms / 1e3, err))
