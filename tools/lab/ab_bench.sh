#!/bin/bash
# alternating A/B of two library builds / settings on the bench workload: bash tools/lab/ab_bench.sh "<env A>" "<env B>" [rounds]
A=$1; B=$2; R=${3:-4}
for i in $(seq $R); do
  for cfg in "$A" "$B"; do
    echo -n "[$cfg] "
    env $cfg python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f M transforms/s, frac %.4f' % (d['value'] / 1e6, d['roofline']['frac']))"
  done
done
