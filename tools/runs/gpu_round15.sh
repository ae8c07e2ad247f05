#!/bin/bash
# GPU run r01x: FFTLog parity after the cheaper non-finite guard + bench line
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_fftlog_gpu.py -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-330 $OUT/bench_$TAG.json
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench2_$TAG.json 2>> $OUT/bench_$TAG.err
cut -c1-330 $OUT/bench2_$TAG.json
