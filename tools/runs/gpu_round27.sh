#!/bin/bash
# GPU run r02s: compute-sanitizer over the kernels touched in this session (memcheck on the parity tests, racecheck on a stream-kernel case)
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fftlog_gpu.py tests/test_eh.py -m gpu -q -x -k "persistent or non_finite or seeded or cuda_generator or multipoles" > $OUT/memcheck_fftlog_$TAG.log 2>&1
echo "memcheck fftlog rc=$?" | tee -a $OUT/memcheck_fftlog_$TAG.log
tail -n 6 $OUT/memcheck_fftlog_$TAG.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_spline_wallish_gpu.py tests/test_interp2d.py -m gpu -q -x > $OUT/memcheck_spline_$TAG.log 2>&1
echo "memcheck spline/wallish rc=$?" | tee -a $OUT/memcheck_spline_$TAG.log
tail -n 6 $OUT/memcheck_spline_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_fftlog_gpu.py -m gpu -q -x -k "persistent and 2001" > $OUT/racecheck_stream_$TAG.log 2>&1
echo "racecheck stream rc=$?" | tee -a $OUT/racecheck_stream_$TAG.log
tail -n 12 $OUT/racecheck_stream_$TAG.log
