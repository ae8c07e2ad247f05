#!/bin/bash
# GPU run r01q: state check after restore (smoke, GPU parity suite, quick bench) + Wallish2018 launch list and ncu captures
TAG=${1:-r01q}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -4 $OUT/pytest_$TAG.log $OUT/smoke_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_wallish_$TAG.csv \
    python tools/bench_extra.py --quick > $OUT/ncu_launch_wallish_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wallish -c 3 -f -o $OUT/prof_wallish_$TAG \
    python tools/bench_extra.py --quick > $OUT/ncu_full_wallish_$TAG.log 2>&1
ls $OUT | grep $TAG
