#!/bin/bash
# GPU run r02c: spline evaluation with per-query pre-kernel: GPU suite + secondary benchmarks
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
timeout 900 python tools/bench_extra.py > $OUT/extra_$TAG.json 2> $OUT/extra_$TAG.err; cat $OUT/extra_$TAG.json; tail -n 3 $OUT/extra_$TAG.err
