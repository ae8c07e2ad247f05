#!/bin/bash
# GPU run r01z (2 GPUs): whole GPU suite on GPU 0, then the bench under torchrun at N = 2 (ours and the reference arm)
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -n 4 $OUT/pytest_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err
echo "bench n2 rc=$?"; cut -c1-700 $OUT/bench_n2_$TAG.json; tail -n 3 $OUT/bench_n2_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2_$TAG.json 2>> $OUT/bench_n2_$TAG.err
echo "ref n2 rc=$?"; cut -c1-300 $OUT/bench_ref_n2_$TAG.json
timeout 300 python tools/bench_extra.py --quick > $OUT/extra_quick_$TAG.json 2>/dev/null; cat $OUT/extra_quick_$TAG.json
