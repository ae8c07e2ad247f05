"""Latency of one host-array call (configs[0]: a single PowerToCorrelation, nk = 1024, ell = 0) and of small batches; the reference's numpy engine beside it."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation
from oracle import fftlog_oracle as O
k = np.geomspace(1e-5, 1e2, 1024)
pk = S.eh_pk(k)
f = PowerToCorrelation(k)
ref = O.execute(O.plan_power_to_correlation(k), pk)[1]
out = f(pk)[1]
print('max |diff| / max |ref| =', float(np.max(np.abs(out - ref)) / np.max(np.abs(ref))))
for B in [1, 4, 16]:
    x = np.tile(pk, (B, 1)) if B > 1 else pk
    f(x)
    t0 = time.perf_counter()
    for _ in range(500): f(x)
    t = (time.perf_counter() - t0) / 500
    plan = O.plan_power_to_correlation(k)
    t0 = time.perf_counter()
    for _ in range(200): O.execute(plan, x)
    tc = (time.perf_counter() - t0) / 200
    print('B = {:2d}: {:.1f} us per call on the GPU path, {:.1f} us for the numpy restatement of the reference'.format(B, t * 1e6, tc * 1e6))
# where the time goes: the bare C call (ctypes, pointers prepared) against the Python wrapper around it
import ctypes
from cosmoprimo_b200 import _lib
lib = _lib.load()
plan = f._device_plan(0)
x = np.ascontiguousarray(pk)
out = np.empty((1, 1, 1024))
args = (plan.handle, x.ctypes.data, 1, 0, 0, 0., 0, 0., 0, out.ctypes.data, 0, 0, None)
lib.cpf_fftlog(*args)
t0 = time.perf_counter()
for _ in range(2000): lib.cpf_fftlog(*args)
print('bare cpf_fftlog (host buffers, B = 1): {:.1f} us'.format((time.perf_counter() - t0) / 2000 * 1e6))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(2000): f(pk)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
