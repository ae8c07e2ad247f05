#!/bin/bash
# GPU run r02u: ping-pong kernel with dynamic scheduling: parity, racecheck, A/B at nk = 1024 / 512
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x -k "persistent and 12001" > $OUT/racecheck_ppdyn_$TAG.log 2>&1
echo "racecheck pp dynamic rc=$?"; tail -n 3 $OUT/racecheck_ppdyn_$TAG.log
python - <<'PY'
import numpy as np, torch, os, json
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import TophatVariance
def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
res = {}
for n in (1024, 512):
    k = np.geomspace(1e-5, 1e2, n)
    base = S.eh_pk(k, S.lhs_cosmologies(1000, seed=42))
    fun = torch.from_numpy(np.repeat(base, 100, axis=0)).cuda()
    tv = TophatVariance(k)
    for rep in range(2):
        for dyn in ('1', '0'):
            os.environ['CPF_STREAM_DYNAMIC'] = dyn
            res['n%d_dyn%s_%d' % (n, dyn, rep)] = 100000 / timed(lambda: tv(fun)) / 1e6
    os.environ.pop('CPF_STREAM_DYNAMIC')
print(json.dumps(res))
PY
