#!/bin/bash
# GPU run r01e: ncu --set full of the ping-pong kernel (large batch and bench config), staged vs zero-copy host path,
# bench with the per-pair and the ping-pong kernel, GPU test-suite.
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log 2>&1
./tools/lab/pp_driver 20 pp0 > $OUT/pp_driver_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_pp -s 2 -c 1 -f -o $OUT/prof_pp0_big_$TAG \
    ./tools/lab/pp_driver 3 pp0 2048 1 100000 > $OUT/ncu_pp0_big_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftlog_pp -s 2 -c 1 -f -o $OUT/prof_pp0_bench_$TAG \
    ./tools/lab/pp_driver 3 pp0 2048 3 4096 > $OUT/ncu_pp0_bench_$TAG.log 2>&1
timeout 300 python tools/lab/e2e_probe.py > $OUT/e2e_probe_staged_$TAG.log 2>&1
CPF_HOST_PATH=zerocopy timeout 300 python tools/lab/e2e_probe.py > $OUT/e2e_probe_zerocopy_$TAG.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
CPF_FFTLOG_KERNEL=pp0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_pp0_$TAG.json 2> $OUT/bench_pp0_$TAG.err
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
cat $OUT/pp_driver_$TAG.log $OUT/e2e_probe_staged_$TAG.log $OUT/e2e_probe_zerocopy_$TAG.log
cut -c1-700 $OUT/bench_$TAG.json; cut -c1-300 $OUT/bench_pp0_$TAG.json
