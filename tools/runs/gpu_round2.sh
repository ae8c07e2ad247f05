#!/bin/bash
# GPU run: full parity suite with both FFTLog kernels, A/B bench, ncu of the persistent kernel and the Wallish kernel
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
CPF_FFTLOG_PERSISTENT=0 timeout 600 python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x > $OUT/pytest_v1_$TAG.log 2>&1; echo "pytest v1 rc=$?" >> $OUT/pytest_v1_$TAG.log
tail -3 $OUT/pytest_v1_$TAG.log
CPF_FFTLOG_PERSISTENT=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_v1_$TAG.json 2> $OUT/bench_v1_$TAG.err
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-400 $OUT/bench_v1_$TAG.json; cat $OUT/bench_$TAG.json
timeout 300 python tools/bench_extra.py > $OUT/extra_$TAG.json 2> $OUT/extra_$TAG.err; cat $OUT/extra_$TAG.json
if [ -z "$SKIP_NCU" ]; then
CPF_BENCH_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
CPF_BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fftlog_persistent -s 4 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wallish_fused -c 1 -f -o $OUT/prof_wallish_$TAG \
    python tools/bench_extra.py --quick > $OUT/ncu_wallish_$TAG.log 2>&1
fi
ls -la $OUT | tail -20
