"""Where does the host-buffer (e2e) call spend its time?  Run on the GPU box: python tools/lab/e2e_probe.py"""
import os
import sys
import time
import ctypes

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
from cosmoprimo_b200 import _lib, _buffers
from cosmoprimo_b200.fftlog import PowerToCorrelation
import bench


def timeit(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print('{:<60s} best {:8.3f} ms   median {:8.3f} ms'.format(name, 1e3 * min(ts), 1e3 * float(np.median(ts))))
    return min(ts)


def main():
    k, fun = bench.make_inputs(bench.NCOSMO, seed=42)
    fftlog = PowerToCorrelation(k, ell=bench.ELLS, engine='cuda', device=0)
    h_fun = torch.from_numpy(fun).pin_memory().numpy()
    nbytes = fun.nbytes
    print('batch bytes in = out = {:.1f} MB'.format(nbytes / 1e6))
    timeit('public API, pinned numpy in -> numpy out', lambda: fftlog(h_fun))
    timeit('public API, pageable numpy in', lambda: fftlog(fun))
    timeit('_host_empty (pinned result buffer) alone', lambda: _buffers._host_empty(fun.shape, 'f8'))
    timeit('np.empty + first touch', lambda: np.empty(fun.shape).fill(0.))
    # raw C call with preallocated pinned buffers
    lib = _lib.load()
    fftlog(h_fun)
    dplan = fftlog._device_plan(0)
    out = torch.empty(fun.shape, dtype=torch.float64, pin_memory=True).numpy()
    ml, vl, mr, vr = 0, 0., 0, 0.

    def raw():
        rc = lib.cpf_fftlog(dplan.handle, h_fun.ctypes.data, fun.shape[0], 1, ml, vl, mr, vr, 0, out.ctypes.data, 0, 0, None)
        assert rc == 0
    t = timeit('cpf_fftlog, preallocated pinned in/out', raw)
    print('  -> {:.2f} M transforms/s, {:.1f} GB/s each way'.format(fun.shape[0] * 3 / t / 1e6, nbytes / t / 1e9))
    d_in = torch.from_numpy(fun).cuda()
    d_out = torch.empty_like(d_in)
    h_in_t = torch.from_numpy(h_fun)
    h_out_t = torch.from_numpy(out)
    timeit('torch H2D copy_ of the batch (pinned)', lambda: d_in.copy_(h_in_t, non_blocking=True))
    timeit('torch D2H copy_ of the batch (pinned)', lambda: h_out_t.copy_(d_out, non_blocking=True))
    timeit('device-resident call', lambda: fftlog(d_in))


if __name__ == '__main__':
    main()
