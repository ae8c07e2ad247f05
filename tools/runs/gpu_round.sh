#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and one full ncu capture of the top kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
if [ -z "$SKIP_NCU" ]; then
CPF_BENCH_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
CPF_BENCH_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fftlog_fast -s 4 -c 2 -f -o $OUT/prof_$TAG \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
fi
ls -la $OUT
