#!/bin/bash
# GPU run r02l: ncu capture of the dynamically scheduled stream kernel (why is it 5x slower?)
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
CPF_BENCH_QUICK=1 CPF_STREAM_DYNAMIC=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fftlog_stream -s 4 -c 1 -f -o $OUT/prof_dyn_$TAG \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_dyn_$TAG.log 2>&1
tail -3 $OUT/ncu_dyn_$TAG.log
