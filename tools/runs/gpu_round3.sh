#!/bin/bash
# GPU run r01c: re-establish the baseline (tests, bench, ncu full with source) + lab microbenchmarks
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log 2>&1
./tools/lab/fft_lab > $OUT/fft_lab_$TAG.log 2>&1
./tools/lab/mio_lab > $OUT/mio_lab_$TAG.log 2>&1
timeout 300 python tools/lab/e2e_probe.py > $OUT/e2e_probe_$TAG.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
CPF_BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fftlog_fast -s 4 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
cat $OUT/fft_lab_$TAG.log $OUT/mio_lab_$TAG.log $OUT/e2e_probe_$TAG.log
cut -c1-600 $OUT/bench_$TAG.json
