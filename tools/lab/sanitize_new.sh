#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (persistent N = 8192 kernel, one-pass log-log spline construction, fused Wallish2018)
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 1100 compute-sanitizer --tool $tool python -m pytest tests -m gpu -x -q -k "(persistent and 4096) or (persistent and 3000) or padlog or (non_finite and pp-4096) or wallish_golden" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" | tail -4
done
