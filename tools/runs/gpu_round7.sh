#!/bin/bash
# GPU run r01h: ablations of the stream kernel (lab build): which resource is the limiter?
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
L=$OUT/stream_abl_$TAG.log
: > $L
for abl in 0 1 2 4 8 12 13 14 15; do
  echo "ABL=$abl" >> $L
  CPF_STREAM_ABL=$abl ./tools/lab/pp_driver_lab 10 stream 2048 1 100000 2>&1 | grep stream >> $L
  CPF_STREAM_ABL=$abl ./tools/lab/pp_driver_lab 20 stream 2048 3 4096 2>&1 | grep stream >> $L
done
cat $L
