#!/bin/bash
# GPU run r02y: Wallish2018 fused kernel with the pivots in shared memory and unrolled eliminations
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_spline_wallish_gpu.py tests/test_interp2d.py -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 3 $OUT/pytest_$TAG.log
for i in 1 2; do timeout 300 python tools/bench_extra.py --quick; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_wallish_$TAG.csv \
    python tools/bench_extra.py --quick > $OUT/ncu_launch_wallish_$TAG.log 2>&1
grep "wallish_fused" $OUT/launches_wallish_$TAG.csv | awk -F'","' '{print $NF}' | head -8
