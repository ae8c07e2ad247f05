#!/bin/bash
# GPU run r02a: host path variants (mixed = direct output writes), GPU suite with the new tests
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ./tools/lab/e2e_lab > $OUT/e2e_lab_$TAG.log 2>&1; cat $OUT/e2e_lab_$TAG.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
CPF_HOST_PATH=mixed timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_mixed_$TAG.json 2> $OUT/bench_$TAG.err
python -c "import json; d=json.load(open('$OUT/bench_mixed_$TAG.json')); print('mixed', d['value'], d['e2e'])"
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_staged_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "import json; d=json.load(open('$OUT/bench_staged_$TAG.json')); print('staged', d['value'], d['e2e'])"
timeout 300 python tools/bench_extra.py --quick > $OUT/extra_quick_$TAG.json 2>/dev/null; cat $OUT/extra_quick_$TAG.json
