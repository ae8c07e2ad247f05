#!/bin/bash
# GPU run r01r: sigma(r) by windowed row splines; Wallish2018 with packed layouts, parallel gap solve, prefetching spline solve
TAG=${1:-r01r}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 6 $OUT/pytest_$TAG.log
timeout 600 python tools/bench_extra.py > $OUT/extra_$TAG.json 2> $OUT/extra_$TAG.err; cat $OUT/extra_$TAG.json; tail -n 5 $OUT/extra_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_wallish_$TAG.csv \
    python tools/bench_extra.py --quick > $OUT/ncu_launch_wallish_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wallish_fused|spline_solve" -c 2 -f -o $OUT/prof_wallish_$TAG \
    python tools/bench_extra.py --quick > $OUT/ncu_full_wallish_$TAG.log 2>&1
ls $OUT | grep $TAG
