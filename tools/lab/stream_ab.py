"""A/B of stream-kernel variants in one process per setting on the same box: bench workload (12288 transforms / launch), 200 launches."""
import os, subprocess, sys
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import bench
from cosmoprimo_b200.fftlog import PowerToCorrelation
k, fun = bench.make_inputs(4096, 42)
d = torch.from_numpy(fun).cuda()
obj = PowerToCorrelation(k, ell=[0, 2, 4])
for _ in range(300): out = obj(d)[1]
torch.cuda.synchronize()
best = 0
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): out = obj(d)[1]
    e1.record(); torch.cuda.synchronize()
    best = max(best, 12288 * 200 / (e0.elapsed_time(e1) * 1e-3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): out = obj(d)[1]
e1.record(); torch.cuda.synchronize()
print('best of 5 x 200 launches: %%.2f M transforms/s ; 20 launches: %%.2f M/s' %% (best / 1e6, 12288 * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e6))
'''
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
settings = [s for s in sys.argv[1:]] or ['', 'CPF_STREAM_FINE=0']
for rnd in range(2):
    for st in settings:
        env = dict(os.environ)
        for kv in st.split(','):
            if kv: env[kv.split('=')[0]] = kv.split('=')[1]
        res = subprocess.run([sys.executable, '-c', CODE % root], env=env, capture_output=True, text=True)
        print('[%s] %s %s' % (st or 'default', res.stdout.strip(), res.stderr[-300:].strip()), flush=True)
