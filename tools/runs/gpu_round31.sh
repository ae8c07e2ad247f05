#!/bin/bash
# GPU run r02x: mirrored reference tests (sigma(r) vs quadrature, NaN tables, 2-D bounds contract)
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_interp2d.py -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 30 $OUT/pytest_$TAG.log
