#!/bin/bash
# GPU run r03f: ncu capture of the Wallish2018 fused kernel at the end of the round
TAG=${1:-r03f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wallish_fused" -c 1 -f -o $OUT/prof_wallish_$TAG \
    python tools/bench_extra.py --quick > $OUT/ncu_full_wallish_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_wallish_$TAG.csv \
    python tools/bench_extra.py --quick > $OUT/ncu_launch_wallish_$TAG.log 2>&1
ls -la $OUT | grep $TAG
