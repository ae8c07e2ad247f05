"""The bench workload (12 288 transforms per launch, stream kernel) timed over 5 x 200 launches: best and median M transforms/s."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from cosmoprimo_b200.fftlog import PowerToCorrelation
k, fun = bench.make_inputs(4096, 42)
d = torch.from_numpy(fun).cuda()
obj = PowerToCorrelation(k, ell=[0, 2, 4])
keep = [obj(d)[1] for _ in range(4)]
for i in range(300): keep[i % 4] = obj(d)[1]
torch.cuda.synchronize()
rates = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): keep[i % 4] = obj(d)[1]
    e1.record(); torch.cuda.synchronize()
    rates.append(12288 * 200 / (e0.elapsed_time(e1) * 1e-3) / 1e6)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(20): keep[i % 4] = obj(d)[1]
e1.record(); torch.cuda.synchronize()
print('200 launches: best %.2f median %.2f M transforms/s ; 20 launches: %.2f' % (max(rates), sorted(rates)[2], 12288 * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e6))
