#!/bin/bash
# compute-sanitizer over the kernels added or changed in round 2 (persistent N = 8192 kernel, one-pass log-log spline construction, chunked
# spline evaluation kernels, fused Wallish2018 with its rows entry, stream kernel)
SEL="(persistent and 4096) or (persistent and 3000) or padlog or (non_finite and pp-4096) or wallish_golden or rows_entry or spline_golden or spline_device or (persistent and stream-2048-601)"
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 1100 compute-sanitizer --tool $tool python -m pytest tests -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" | tail -4
done
