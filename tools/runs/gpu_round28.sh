#!/bin/bash
# GPU run r02t: racecheck on the ping-pong kernel (TMA staging), the per-pair kernel and the Wallish2018 / spline kernels; initcheck on the FFTLog suite
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x -k "(persistent and pp) or non_finite" > $OUT/racecheck_pp_$TAG.log 2>&1
echo "racecheck pp/fast rc=$?" | tee -a $OUT/racecheck_pp_$TAG.log
tail -n 4 $OUT/racecheck_pp_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_spline_wallish_gpu.py -m gpu -q -x -k "wallish_golden or eval_rows or dst" > $OUT/racecheck_wallish_$TAG.log 2>&1
echo "racecheck wallish rc=$?" | tee -a $OUT/racecheck_wallish_$TAG.log
tail -n 4 $OUT/racecheck_wallish_$TAG.log
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x -k "persistent or non_finite" > $OUT/initcheck_fftlog_$TAG.log 2>&1
echo "initcheck fftlog rc=$?" | tee -a $OUT/initcheck_fftlog_$TAG.log
tail -n 4 $OUT/initcheck_fftlog_$TAG.log
