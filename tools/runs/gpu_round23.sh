#!/bin/bash
# GPU run r02h (lab build): per-CTA time line of the stream kernel (TMA staging) in the bench configuration
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
CPF_STREAM_DBG=2 ./tools/lab/pp_driver_lab 3 stream 2048 3 4096 2>&1 | tail -24 > $OUT/stream_timeline_$TAG.log
cat $OUT/stream_timeline_$TAG.log
