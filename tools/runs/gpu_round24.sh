#!/bin/bash
# GPU run r02k: ping-pong kernel with TMA staging, stream kernel with dynamic pair scheduling: parity + A/B
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 5 $OUT/pytest_$TAG.log
for dyn in 1 0 1 0; do  # CPF_STREAM_DYNAMIC is opt-in: 1 = tickets, 0 = static split
  CPF_STREAM_DYNAMIC=$dyn timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_dyn${dyn}_$TAG.json 2>> $OUT/bench_$TAG.err
  python -c "import json; d=json.load(open('$OUT/bench_dyn${dyn}_$TAG.json')); print('dynamic $dyn', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
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
for n in (1024, 512, 2048):
    k = np.geomspace(1e-5, 1e2, n)
    base = S.eh_pk(k, S.lhs_cosmologies(1000, seed=42))
    fun = torch.from_numpy(np.repeat(base, 100, axis=0)).cuda()
    tv = TophatVariance(k)
    for tma in ('1', '0'):
        os.environ['CPF_STREAM_TMA'] = tma
        res['n%d_tma%s' % (n, tma)] = 100000 / timed(lambda: tv(fun)) / 1e6
    os.environ.pop('CPF_STREAM_TMA')
    for dyn in ('1', '0'):
        os.environ['CPF_STREAM_DYNAMIC'] = dyn
        res['n%d_dyn%s' % (n, dyn)] = 100000 / timed(lambda: tv(fun)) / 1e6
    os.environ.pop('CPF_STREAM_DYNAMIC')
print(json.dumps(res))
PY
