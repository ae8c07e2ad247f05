#!/bin/bash
# GPU run r01n: stream kernel with dynamic pair scheduling / coalesced tables: A/B, time line, static vs dynamic
TAG=${1:-r01n}
OUT=gpurun_out
mkdir -p $OUT
./tools/lab/pp_driver 20 stream > $OUT/pp_driver_$TAG.log 2>&1
echo "--- static partition" >> $OUT/pp_driver_$TAG.log
CPF_STREAM_STATIC=1 ./tools/lab/pp_driver 20 stream 2048 3 4096 | grep stream >> $OUT/pp_driver_$TAG.log 2>&1
CPF_STREAM_STATIC=1 ./tools/lab/pp_driver 10 stream 2048 1 100000 | grep stream >> $OUT/pp_driver_$TAG.log 2>&1
cat $OUT/pp_driver_$TAG.log
CPF_STREAM_DBG=2 ./tools/lab/pp_driver_lab 2 stream 2048 3 4096 2>&1 | tail -19 > $OUT/stream_timeline_$TAG.log
cat $OUT/stream_timeline_$TAG.log
