#!/bin/bash
# GPU run r02q: EH generator with the coefficient kernel split off: parity + rows/s
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_eh.py tests/test_interp2d.py -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 3 $OUT/pytest_$TAG.log
python - <<'PY'
import numpy as np, torch, json
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.eisenstein_hu import EisensteinHu
def timed(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
k = np.geomspace(1e-5, 1e2, 2048)
res = {}
for B in (4096, 10000, 100000):
    par = S.lhs_cosmologies(B, seed=42)
    eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
    zz = np.linspace(0., 3., B)
    res['single_z_B%d_rows_per_s' % B] = B / timed(lambda: eh.pk(k, z=zz))
    res['kaiser_B%d_cosmologies_per_s' % B] = B / timed(lambda: eh.pk(k, z=zz, kaiser=True))
par = S.lhs_cosmologies(10000, seed=42)
eh = EisensteinHu(par['h'], par['omega_b'], par['omega_cdm'], par['n_s'], logA=par['logA'])
zg = np.linspace(0., 3., 100)[None, :]
res['zgrid_rows_per_s'] = 1e6 / timed(lambda: eh.pk(k, z=zg), reps=3, warm=1)
print(json.dumps(res))
PY
