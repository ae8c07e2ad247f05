#!/bin/bash
# GPU run r02f (lab build): phase offset between the two groups of the stream kernel
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
L=$OUT/stream_skew_$TAG.log
: > $L
for skew in 0 900 1800 2700 3600 5400 0 1800; do
  echo "--- CPF_STREAM_SKEW_NS=$skew" >> $L
  CPF_STREAM_SKEW_NS=$skew timeout 120 ./tools/lab/pp_driver_lab 30 stream 2048 3 4096 2>&1 | grep "stream " >> $L
  CPF_STREAM_SKEW_NS=$skew timeout 120 ./tools/lab/pp_driver_lab 10 stream 2048 1 100000 2>&1 | grep "stream " >> $L
done
cat $L
