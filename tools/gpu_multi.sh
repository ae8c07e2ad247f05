#!/bin/bash
# Multi-GPU record (gpurun --gpus N): topology, the default bench and the configs[4] round-trip workload under torchrun, the NCCL gather test.
#   usage: bash tools/gpu_multi.sh <tag> <ngpus> [steps]
TAG=${1:-r00}; N=${2:-2}; STEPS=${3:-20}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${TAG}.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $RUN bench.py --gpus $N --steps $STEPS --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench rc=$?"; cut -c1-2500 $OUT/bench_n${N}_$TAG.json; tail -3 $OUT/bench_n${N}_$TAG.err
CPF_NO_NUMA_BIND=1 timeout 900 $RUN bench.py --gpus $N --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/bench_nobind_n${N}_$TAG.json 2> $OUT/bench_nobind_n${N}_$TAG.err; echo "bench (no NUMA binding) rc=$?"; python - <<PY
import json
for name in ['bench_n${N}_$TAG', 'bench_nobind_n${N}_$TAG']:
    try:
        d = json.load(open('$OUT/' + name + '.json'))
        print(name, 'value %.1f M/s' % (d['value'] / 1e6), 'e2e %.2f M/s' % (d['e2e']['value'] / 1e6), {k: round(v['value'] / 1e6, 2) for k, v in d['e2e'].get('variants', {}).items()}, d['e2e'].get('numa_binding_rank0'))
    except Exception as e:
        print(name, 'unreadable', e)
PY
timeout 900 $RUN bench.py --gpus $N --workload roundtrip --steps 5 --warmup 3 > $OUT/roundtrip_n${N}_$TAG.json 2> $OUT/roundtrip_n${N}_$TAG.err; echo "roundtrip rc=$?"; cat $OUT/roundtrip_n${N}_$TAG.json; tail -3 $OUT/roundtrip_n${N}_$TAG.err
timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q > $OUT/pytest_dist_n${N}_$TAG.log 2>&1; tail -3 $OUT/pytest_dist_n${N}_$TAG.log
ls $OUT | grep $TAG
