#!/bin/bash
# GPU run r01u: 2-D interpolators, sigma stage timings, full-size configs 3 and 4
TAG=${1:-r01u}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 6 $OUT/pytest_$TAG.log
timeout 300 python tools/lab/sigma_stages.py > $OUT/sigma_stages_$TAG.json 2> $OUT/sigma_stages_$TAG.err; cat $OUT/sigma_stages_$TAG.json; tail -n 3 $OUT/sigma_stages_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches_sigma_$TAG.csv \
    python tools/lab/sigma_stages.py > $OUT/ncu_launch_sigma_$TAG.log 2>&1
timeout 900 python tools/bench_extra.py > $OUT/extra_$TAG.json 2> $OUT/extra_$TAG.err; cat $OUT/extra_$TAG.json; tail -n 5 $OUT/extra_$TAG.err
ls $OUT | grep $TAG
