#!/bin/bash
# GPU run r01m: time line of the stream kernel in the bench configuration (lab build), ncu of the production build
TAG=${1:-r01m}
OUT=gpurun_out
mkdir -p $OUT
CPF_STREAM_DBG=2 ./tools/lab/pp_driver_lab 2 stream 2048 3 4096 > $OUT/stream_timeline_$TAG.log 2>&1
tail -40 $OUT/stream_timeline_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream -s 2 -c 1 -f -o $OUT/prof_stream_bench_$TAG \
    ./tools/lab/pp_driver 3 stream 2048 3 4096 > $OUT/ncu_stream_bench_$TAG.log 2>&1
