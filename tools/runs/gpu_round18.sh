#!/bin/bash
# GPU run r02b: stream-kernel prefetch variants (lab build): L1 prefetch of the next pair, deeper L2 prefetch
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
L=$OUT/stream_prefetch_$TAG.log
: > $L
for abl in 0 16 32 48 0 16; do
  echo "--- CPF_STREAM_ABL=$abl bench config (n 2048, P 3, 4096 rows)" >> $L
  CPF_STREAM_ABL=$abl ./tools/lab/pp_driver_lab 30 stream 2048 3 4096 2>&1 | grep -i "stream\|us" | tail -3 >> $L
  echo "--- CPF_STREAM_ABL=$abl n 2048, P 1, 100000 rows" >> $L
  CPF_STREAM_ABL=$abl ./tools/lab/pp_driver_lab 10 stream 2048 1 100000 2>&1 | grep -i "stream\|us" | tail -3 >> $L
done
cat $L
