#!/bin/bash
# GPU run r03k (N GPUs of one box): the bench under torchrun at N = $1, ours and the reference arm
N=${1:-8}
TAG=${2:-r03k}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
echo "bench n$N rc=$?"; cut -c1-400 $OUT/bench_n${N}_$TAG.json; tail -n 3 $OUT/bench_n${N}_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n${N}_$TAG.json 2>> $OUT/bench_n${N}_$TAG.err
echo "ref n$N rc=$?"; cut -c1-300 $OUT/bench_ref_n${N}_$TAG.json
