#!/bin/bash
# GPU run r02g: stream kernel launched with programmatic stream serialisation (PDL) vs plain launches
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 3 $OUT/pytest_$TAG.log
for pdl in 1 0 1 0; do
  CPF_STREAM_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_pdl${pdl}_$TAG.json 2>> $OUT/bench_$TAG.err
  python -c "import json; d=json.load(open('$OUT/bench_pdl${pdl}_$TAG.json')); print('pdl $pdl', d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
timeout 600 python tools/bench_extra.py --quick > /dev/null 2>&1
python - <<'PY'
import numpy as np, torch, time
from cosmoprimo_b200 import synthetic as S
from cosmoprimo_b200.fftlog import PowerToCorrelation, CorrelationToPower
# dependent back-to-back kernels: round trip P -> xi -> P must still be exact (the second kernel reads what the first wrote)
k = np.geomspace(1e-5, 1e2, 2048)
pk = torch.from_numpy(S.eh_pk(k, S.lhs_cosmologies(3000, seed=1))).cuda()
p2x = PowerToCorrelation(k); s, xi = p2x(pk); x2p = CorrelationToPower(s)
ref = x2p(xi)[1].clone(); torch.cuda.synchronize()
bad = 0
for it in range(50):
    out = x2p(p2x(pk)[1])[1]
    bad += int(not torch.equal(out, ref))
torch.cuda.synchronize()
print('round trip with dependent launches: mismatches', bad, 'of 50; max rel dev from input', float(((out / pk - 1).abs())[:, 600:1400].max()))
PY
