#!/bin/bash
# GPU run r01f: stream kernel A/B against the per-pair and ping-pong kernels, ncu --set full of the stream kernel
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
./tools/lab/pp_driver 20 stream > $OUT/pp_driver_stream_$TAG.log 2>&1
./tools/lab/pp_driver 20 pp0 2048 1 100000 >> $OUT/pp_driver_stream_$TAG.log 2>&1
cat $OUT/pp_driver_stream_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream -s 2 -c 1 -f -o $OUT/prof_stream_big_$TAG \
    ./tools/lab/pp_driver 3 stream 2048 1 100000 > $OUT/ncu_stream_big_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream -s 2 -c 1 -f -o $OUT/prof_stream_bench_$TAG \
    ./tools/lab/pp_driver 3 stream 2048 3 4096 > $OUT/ncu_stream_bench_$TAG.log 2>&1
CPF_FFTLOG_KERNEL=stream timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_stream_$TAG.json 2> $OUT/bench_stream_$TAG.err
cut -c1-300 $OUT/bench_stream_$TAG.json
