#!/bin/bash
# GPU run r02d: stream kernel with TMA-staged input rows vs direct loads (A/B through the C ABI), parity, bench
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
L=$OUT/stream_tma_$TAG.log
: > $L
for tma in 0 1 0 1; do
  echo "--- CPF_STREAM_TMA=$tma" >> $L
  CPF_STREAM_TMA=$tma timeout 120 ./tools/lab/pp_driver 30 stream 2048 3 4096 2>&1 | grep "stream" >> $L
  CPF_STREAM_TMA=$tma timeout 120 ./tools/lab/pp_driver 10 stream 2048 1 100000 2>&1 | grep "stream" >> $L
  CPF_STREAM_TMA=$tma timeout 120 ./tools/lab/pp_driver 10 stream 2048 3 4097 2>&1 | grep "stream" >> $L
done
cat $L
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python -c "import json; d=json.load(open('$OUT/bench_$TAG.json')); print('bench', d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity'])"
CPF_STREAM_TMA=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_notma_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "import json; d=json.load(open('$OUT/bench_notma_$TAG.json')); print('bench no tma', d['value'], d['roofline']['frac'])"
