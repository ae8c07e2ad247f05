#!/bin/bash
# GPU run r01i: stream2 variants (2/3 groups of 128 threads, interleaved/sequential column sets)
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
L=$OUT/pp_driver_variants_$TAG.log
: > $L
for k in stream stream2 stream3 stream4 stream5; do
  ./tools/lab/pp_driver 10 $k 2048 1 100000 2>&1 | grep "$k " >> $L
  ./tools/lab/pp_driver 20 $k 2048 3 4096 2>&1 | grep "$k " >> $L
done
./tools/lab/pp_driver 10 stream4 >> $L 2>&1
cat $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream2 -s 2 -c 1 -f -o $OUT/prof_stream4_big_$TAG \
    ./tools/lab/pp_driver 3 stream4 2048 1 100000 > $OUT/ncu_stream4_big_$TAG.log 2>&1
