#!/bin/bash
# GPU run r02w: bench.py --secondary (other BASELINE configs with the oracle's CPU timing beside them)
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --secondary > $OUT/secondary_$TAG.json 2> $OUT/secondary_$TAG.err
echo "rc=$?"; cat $OUT/secondary_$TAG.json; tail -n 5 $OUT/secondary_$TAG.err
