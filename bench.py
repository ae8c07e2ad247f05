"""
Benchmark of the FFTLog hot path (BASELINE.json metric: FFTLog transforms/s, fp64, nk=2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[1]: batched P(k) -> xi multipoles ell=0,2,4, nk=2048 (padded N=4096), 4096 synthetic
Eisenstein-Hu cosmologies (Latin hypercube, seed 42 + rank) => 12288 transforms per step and per GPU.
One step = one pass of the fused kernel over that batch.  Input + output (2 x 201 MB) exceed the 126 MB L2, so every
step streams from HBM (no explicit flush needed).

`value`  : transforms/s with inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks.
`e2e`    : the same through the public API with pinned HOST arrays in and out (H2D + D2H inside the timed region).
`roofline`: achieved vs the binding roofline of BASELINE.md §3, max(bytes/BW_HBM, flops/F_fp64); HBM peak from
            MEASURED_PEAKS.json, fp64 peak measured here by a DFMA microbenchmark (cpf_measure_fp64_peak) because
            MEASURED_PEAKS.json has no fp64 entry.
`cpu_baseline`: the numpy restatement of the reference path (oracle/, kind "port"; it calls numpy.fft exactly as
            cosmoprimo's NumpyFFTEngine does) on all host cores, bounded sample, rank 0 at N=1 only.
`--impl reference` times that CPU path alone with the same JSON schema.
"""

import os
import sys
import json
import time
import argparse
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NK = 2048
NCOSMO = 4096
ELLS = [0, 2, 4]
METRIC = 'fftlog_transforms_per_sec_fp64_nk2048'
UNIT = 'transforms/s'
# algorithmic work per transform, SURVEY.md §8(d) / BASELINE.md §3
BYTES_PER_TRANSFORM = 16 * NK
FLOPS_PER_TRANSFORM = 2 * 2.5 * (2 * NK) * np.log2(2 * NK) + 2 * NK + 6 * (NK + 1) + NK     # = 264198
CPU_SAMPLE_COSMO = 256     # cosmologies per worker and repetition in the CPU legs
QUICK = bool(os.environ.get('CPF_BENCH_QUICK'))   # profiler runs: no time-based warm-up / sustain loops


def make_inputs(ncosmo, seed):
    from cosmoprimo_b200 import synthetic
    k = np.geomspace(1e-5, 1e2, NK)
    pk = synthetic.eh_pk(k, synthetic.lhs_cosmologies(ncosmo, seed=seed))
    return k, synthetic.kaiser_multipoles(pk, np.full(ncosmo, 0.76))


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference path).  Workers are spawned processes that import numpy + oracle only.
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(root, seed_base):
    os.environ['OMP_NUM_THREADS'] = '1'
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import fftlog_oracle as O
    k, fun = make_inputs(CPU_SAMPLE_COSMO, seed_base + os.getpid() % 1000)
    _W['O'], _W['plan'], _W['fun'] = O, O.plan_power_to_correlation(k, ell=ELLS), fun


def _cpu_step(reps):
    O, plan, fun = _W['O'], _W['plan'], _W['fun']
    t0 = time.perf_counter()
    for _ in range(reps):
        O.execute(plan, fun)
    return reps * fun.shape[0] * fun.shape[1], time.perf_counter() - t0


class CpuPool(object):

    def __init__(self, cores=None):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = ProcessPoolExecutor(max_workers=self.cores, mp_context=mp.get_context('spawn'),
                                        initializer=_cpu_init, initargs=(ROOT, 1000))
        self.step(1)   # start every worker

    def step(self, reps):
        """All workers transform their sample `reps` times concurrently; returns (transforms, wall seconds)."""
        t0 = time.perf_counter()
        res = list(self.pool.map(_cpu_step, [reps] * self.cores))
        return sum(r[0] for r in res), time.perf_counter() - t0

    def close(self):
        self.pool.shutdown()


def cpu_sample_desc(cores, reps):
    return ('oracle port of cosmoprimo numpy engine (numpy {} pocketfft, 1 thread/process): {} processes x {} reps x '
            '({} cosmologies x 3 ell, nk={}) per step').format(np.__version__, cores, reps, CPU_SAMPLE_COSMO, NK)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    reps = 2
    pool = CpuPool()
    for _ in range(max(args.warmup, 1)):
        pool.step(reps)
    total, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        total += pool.step(reps)[0]
    elapsed = time.perf_counter() - t0
    pool.close()
    value = total / elapsed
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / args.steps, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': 'P(k)->xi multipoles ell=0,2,4, nk=2048, synthetic EH cosmologies (BASELINE configs[1])',
                      'batch': '{} cosmologies x 3 ell per worker and rep'.format(CPU_SAMPLE_COSMO), 'nk': NK, 'padded_size': 2 * NK},
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': pool.cores, 'kind': 'port', 'sample': cpu_sample_desc(pool.cores, reps)},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.proc = device, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits', '-lms', '50',
                                          '-i', str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            text = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            text = ''
        sm, smmax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in text.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        # samples under load = the upper half by power draw (the sampler also sees idle gaps around the run)
        order = np.argsort(power)[len(power) // 2:]
        return {'sm_mhz': float(np.median(np.asarray(sm)[order])), 'sm_max_mhz': float(np.max(smmax)),
                'power_w_max': float(np.max(power)), 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    fn = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(fn):
        with open(fn) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650., 'fallback (B200_PROFILING.md 6.65 TB/s)'


def run_ours(args):
    import ctypes
    import torch
    from cosmoprimo_b200 import _lib
    from cosmoprimo_b200.fftlog import PowerToCorrelation

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective: send it to stderr so that stdout carries
        # the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            torch.cuda.set_device(local_rank)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    _lib.require_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # rows are independent: each rank owns its own 4096 cosmologies (weak scaling, no data-path collective)
    k, fun = make_inputs(NCOSMO, seed=42 + rank)
    per_step = fun.shape[0] * fun.shape[1]
    fftlog = PowerToCorrelation(k, ell=ELLS, engine='cuda', device=local_rank)
    d_fun = torch.from_numpy(fun).to(dev)

    def step():
        return fftlog(d_fun)[1]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # warm-up: at least W steps and at least 0.3 s of the same load so that clocks settle
    # (results are kept alive across calls exactly as in the timed loop, so that the caching allocator already owns
    # both 201 MB result blocks: a cudaMalloc inside the timed region would synchronise the device)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        out = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < (0. if QUICK else 0.3):
        for _ in range(20):
            out = step()
        torch.cuda.synchronize()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    # keep the same load running ~1 s longer so that the 50 ms clock sampler sees it (the timed region is only a few ms)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < (0. if QUICK else 1.0):
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    value = world * per_step * args.steps / elapsed

    # end to end through the public API with pinned host arrays (H2D + D2H inside the timed region)
    h_fun = torch.from_numpy(fun).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):   # results are kept alive across calls, as in the timed loop, so that the pinned-buffer cache is warm
        h_out = fftlog(h_fun)[1]
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_out = fftlog(h_fun)[1]
    torch.cuda.synchronize()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * per_step * e2e_steps / e2e_elapsed
    # the host path runs the per-pair kernel on 16 MB chunks, the device path the persistent kernel: same transform, different rounding
    d_out = out.cpu().numpy()
    assert np.max(np.abs(h_out - d_out)) <= 1e-13 * np.max(np.abs(d_out))

    if rank != 0:
        return

    # roofline of the fused kernel: one launch per step
    hbm_gbs, hbm_src = measured_peaks()
    f64 = ctypes.c_double(0.)
    _lib.check(_lib.load().cpf_measure_fp64_peak(local_rank, ctypes.byref(f64)))
    f64_tflops = f64.value / 1e12
    t_launch = elapsed / args.steps
    ach_tflops = per_step * FLOPS_PER_TRANSFORM / t_launch / 1e12
    ach_gbs = per_step * BYTES_PER_TRANSFORM / t_launch / 1e9
    t_fp64, t_hbm = FLOPS_PER_TRANSFORM / (f64_tflops * 1e12), BYTES_PER_TRANSFORM / (hbm_gbs * 1e9)
    if t_fp64 >= t_hbm:
        roof = {'bound': 'fp64', 'achieved': ach_tflops, 'peak': f64_tflops, 'unit': 'TFLOP/s', 'frac': ach_tflops / f64_tflops}
    else:
        roof = {'bound': 'hbm', 'achieved': ach_gbs, 'peak': hbm_gbs, 'unit': 'GB/s', 'frac': ach_gbs / hbm_gbs}
    # DRAM bytes of one launch of the step's kernel from the committed `ncu --set full` capture (profiles/r03i_ncu_stream_kernel.txt)
    roof.update({'traffic': 347029504, 'traffic_unit': 'bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r03i)', 'kernel': 'fftlog_stream_kernel<fullwin, tma-staged rows, dynamic pairs> (CPF_FFTLOG_KERNEL=%s)' % os.environ.get('CPF_FFTLOG_KERNEL', 'auto'), 'launch_ms': 1e3 * t_launch,
                 'algorithmic_flops_per_launch': per_step * FLOPS_PER_TRANSFORM, 'algorithmic_bytes_per_launch': per_step * BYTES_PER_TRANSFORM,
                 'peak_source': 'fp64: DFMA microbenchmark in this run (cpf_measure_fp64_peak); hbm: ' + hbm_src,
                 'hbm': {'achieved': ach_gbs, 'peak': hbm_gbs, 'unit': 'GB/s', 'frac': ach_gbs / hbm_gbs},
                 'roofline_transforms_per_s': 1. / max(t_fp64, t_hbm)})

    result = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
              'ms_per_step': 1e3 * elapsed / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
              'dtype': 'f64', 'data': 'synthetic',
              'config': {'workload': 'P(k)->xi multipoles ell=0,2,4, nk=2048, 4096 synthetic EH cosmologies per GPU (BASELINE configs[1])',
                         'transforms_per_step_per_gpu': per_step, 'nk': NK, 'padded_size': 2 * NK,
                         'l2': 'inputs+outputs (403 MB/step) larger than L2, no flush', 'partition': 'rows split over ranks, no collective'},
              'clocks': clocks,
              'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(fun.nbytes), 'd2h_bytes_per_step': int(h_out.nbytes), 'steps': e2e_steps},
              'gpu_launches': args.steps, 'roofline': roof}

    if world == 1 and not args.no_cpu_baseline:
        # CPU baseline + parity of the GPU result against the oracle on the same sample
        from oracle import fftlog_oracle as O
        ref = O.execute(O.plan_power_to_correlation(k, ell=ELLS), fun[:64])[1]
        post = fftlog.padded_postfactor[:, fftlog.padded_size_out_left:fftlog.padded_size_out_left + NK]
        result['parity'] = {'scale_aware_max_err': float(np.max(O.scale_aware_error(h_out[:64], ref, post))), 'rows': 64 * 3, 'tol': 1e-10}
        pool = CpuPool()
        reps = 2
        pool.step(reps)
        best = 0.
        for _ in range(5):
            n, t = pool.step(reps)
            best = max(best, n / t)
        pool.close()
        single = CpuPool(cores=1)
        n, t = single.step(reps)
        single.close()
        result['cpu_baseline'] = {'value': best, 'unit': UNIT, 'cores': pool.cores, 'kind': 'port', 'sample': cpu_sample_desc(pool.cores, reps),
                                  'value_1core': n / t}
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args):
    """
    `--secondary`: the other BASELINE.json configs on one GPU, each with the oracle port of the reference timed beside it on one
    host core (bounded sample; SURVEY.md section 8d asks for the reference's CPU path next to every GPU number).  One JSON line.
    Not part of the driver's contract: the default invocation is unchanged.
    """
    import torch
    from cosmoprimo_b200 import synthetic as S, _lib
    from cosmoprimo_b200.fftlog import TophatVariance
    from cosmoprimo_b200.interp import spline_eval_rows, Interpolator1D
    from cosmoprimo_b200.interpolator import PowerSpectrumInterpolator1D
    from cosmoprimo_b200.eisenstein_hu import EisensteinHu
    from cosmoprimo_b200.bao_filter import PowerSpectrumBAOFilter
    from oracle import fftlog_oracle as O, spline_oracle as SO, wallish_oracle as WO
    _lib.require_device()

    def gpu_time(fn, reps=5, warm=2):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    def cpu_time(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        return (time.perf_counter() - t0) / reps

    res = {'secondary': True, 'cpu': 'oracle port (numpy {} / scipy), 1 core, bounded samples'.format(np.__version__), 'data': 'synthetic'}
    n = 2048
    k = np.geomspace(1e-5, 1e2, n)
    r = np.linspace(1., 20., 10)
    # config 3: sigma(r, z), 10 000 cosmologies x 100 redshifts, generated, transformed and reduced on the device
    ncosmo, nz = 10000, 100
    par = S.lhs_cosmologies(ncosmo, seed=42)
    eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    zgrid = np.linspace(0., 3., nz)[None, :]
    tv = TophatVariance(k)
    s_grid = tv.y if tv.y.ndim == 1 else tv.y[0]
    t = gpu_time(lambda: spline_eval_rows(s_grid, tv(eh.pk(k, z=zgrid).reshape(ncosmo * nz, n))[1], r)**0.5, reps=3, warm=1)
    sub = {name: val[:2] for name, val in par.items()}
    pk_cpu = (S.eh_pk(k, sub)[:, None, :] * np.ones((1, 128, 1))).reshape(256, n)
    plan = O.plan_tophat_variance(k)

    def sigma_cpu():
        s, var = O.execute(plan, pk_cpu)
        return SO.interpolator1d(s, var.T, assume_sorted=True)(r)**0.5
    tc = cpu_time(sigma_cpu)
    res['config3_sigma_rz'] = {'unit': 'rows/s', 'gpu': ncosmo * nz / t, 'gpu_rows': ncosmo * nz, 'cpu_1core': 256 / tc, 'cpu_rows': 256,
                               'gpu_path': 'EH generator + TophatVariance + row splines at 10 radii, all on the device',
                               'cpu_path': 'TophatVariance (numpy FFTs) + scipy CubicSpline + evaluation at 10 radii, spectra given'}
    tg = gpu_time(lambda: eh.pk(k, z=zgrid), reps=3, warm=1)
    tcg = cpu_time(lambda: S.eh_pk(k, {name: np.repeat(val[:8], 32) for name, val in par.items()}, z=np.tile(np.linspace(0., 3., 32), 8)))
    res['eh_generator'] = {'unit': 'rows/s', 'gpu': ncosmo * nz / tg, 'cpu_1core': 256 / tcg,
                           'cpu_path': 'numpy restatement of the reference engine (cosmoprimo_b200/synthetic.py), one redshift per row'}
    del eh
    torch.cuda.empty_cache()
    # config 4: Wallish2018 over 65 536 spectra (filter only: the two evaluations of the input interpolator are inputs)
    ktab = np.geomspace(1e-5, 1e2, 512)
    base = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=42)).T
    ncols = 65536
    pk = torch.from_numpy(np.tile(base, (1, ncols // 256)) * (1 + 1e-3 * np.arange(ncols) / ncols)).cuda()
    interp = PowerSpectrumInterpolator1D(ktab, pk)
    filt = PowerSpectrumBAOFilter(interp, engine='wallish2018')
    klin = np.linspace(interp.extrap_kmin, 2., 4096)
    pklin, pkout = interp(klin), interp(filt.k)
    lib = _lib.load()
    kl, ko = torch.from_numpy(klin).cuda(), torch.from_numpy(filt.k).cuda()
    out = torch.empty_like(pkout)
    stream = torch.cuda.current_stream().cuda_stream
    tw = gpu_time(lambda: _lib.check(lib.cpf_wallish2018(kl.data_ptr(), pklin.data_ptr(), 4096, ko.data_ptr(), pkout.data_ptr(), filt.k.size, ncols,
                                                         out.data_ptr(), None, 1, 0, stream)), reps=3, warm=1)
    tf = gpu_time(lambda: filt(interp), reps=3, warm=1)
    pl_cpu, po_cpu = pklin[:, :32].cpu().numpy(), pkout[:, :32].cpu().numpy()
    twc = cpu_time(lambda: WO.wallish2018(klin, pl_cpu, filt.k, po_cpu), reps=2)
    res['config4_wallish2018'] = {'unit': 'P(k)/s', 'gpu_filter_only': ncols / tw, 'gpu_with_input_evaluations': ncols / tf, 'gpu_spectra': ncols,
                                  'cpu_1core_filter_only': 32 / twc, 'cpu_spectra': 32,
                                  'cpu_path': 'scipy dst / CubicSpline restatement of Wallish2018PowerSpectrumBAOFilter._compute (per-column Python loop as in the reference)'}
    del pk, interp, filt, pklin, pkout, out
    torch.cuda.empty_cache()
    # spline evaluation of a log-log P(k) table (the step in front of every transform): 540 knots x 4096 spectra -> 2048 wavenumbers
    ktab = np.geomspace(1e-4, 50., 540)
    tab = S.eh_pk(ktab, S.lhs_cosmologies(256, seed=3)).T
    tab_d = torch.from_numpy(np.tile(tab, (1, 16))).cuda()
    kq = np.geomspace(1e-4, 50., n)
    ti = gpu_time(lambda: Interpolator1D(ktab, tab_d, interp_x='log', interp_fun='log', assume_sorted=True).eval_rows(kq))
    tic = cpu_time(lambda: SO.interpolator1d(ktab, tab, interp_x='log', interp_fun='log', assume_sorted=True)(kq))
    res['spline_fit_and_eval'] = {'unit': 'spectra/s', 'gpu': tab_d.shape[1] / ti, 'cpu_1core': tab.shape[1] / tic,
                                  'what': 'natural cubic spline fit (540 knots, log-log) + evaluation at 2048 wavenumbers'}
    print(json.dumps(res))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=200)
    parser.add_argument('--warmup', type=int, default=10)
    parser.add_argument('--impl', type=str, default='ours', choices=['ours', 'reference'])
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--secondary', action='store_true', help='the other BASELINE configs with the oracle timed beside them (one GPU)')
    args = parser.parse_args()
    if args.secondary:
        run_secondary(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
