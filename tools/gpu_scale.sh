#!/bin/bash
# Scaling record on one box (gpurun --gpus 8): the default bench and the configs[4] round trip at N = 1, 2, 4, 8, back to back.
#   usage: bash tools/gpu_scale.sh <tag>
TAG=${1:-r00}; OUT=gpurun_out; mkdir -p $OUT
: > $OUT/scale_$TAG.jsonl; : > $OUT/scale_roundtrip_$TAG.jsonl
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N))"; fi
  timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>>$OUT/scale_$TAG.err | tail -1 >> $OUT/scale_$TAG.jsonl
  timeout 600 $RUN bench.py --gpus $N --workload roundtrip --steps 5 --warmup 3 2>>$OUT/scale_$TAG.err | tail -1 >> $OUT/scale_roundtrip_$TAG.jsonl
done
python - <<PY
import json
for name in ['scale_$TAG', 'scale_roundtrip_$TAG']:
    rows = [json.loads(l) for l in open('$OUT/' + name + '.jsonl') if l.strip().startswith('{')]
    base = rows[0]['value'] / rows[0]['n_gpus'] if rows else 1.
    for d in rows:
        e2e = d.get('e2e', {}).get('value')
        print(name, 'N =', d['n_gpus'], '%.1f M/s' % (d['value'] / 1e6), 'x%.2f' % (d['value'] / base), ('e2e %.2f M/s' % (e2e / 1e6)) if e2e else '', d.get('roofline', {}).get('frac'))
PY
