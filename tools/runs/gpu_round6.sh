#!/bin/bash
# GPU run r01g: stream2 (two column sets per thread) A/B + ncu
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
./tools/lab/pp_driver 20 stream2 > $OUT/pp_driver_stream2_$TAG.log 2>&1
./tools/lab/pp_driver 20 stream 2048 1 100000 >> $OUT/pp_driver_stream2_$TAG.log 2>&1
cat $OUT/pp_driver_stream2_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream2 -s 2 -c 1 -f -o $OUT/prof_stream2_big_$TAG \
    ./tools/lab/pp_driver 3 stream2 2048 1 100000 > $OUT/ncu_stream2_big_$TAG.log 2>&1
CPF_FFTLOG_KERNEL=stream2 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_stream2_$TAG.json 2> $OUT/bench_stream2_$TAG.err
cut -c1-300 $OUT/bench_stream2_$TAG.json
