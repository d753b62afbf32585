#!/bin/bash
# multi-GPU bench (run under: gpurun --gpus N -- bash tools/gpu_multi.sh <tag> N)
TAG=$1; N=$2; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "rc=$?"
tail -3 $OUT/${TAG}_bench_n$N.err; cat $OUT/${TAG}_bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $OUT/${TAG}_ref_n$N.json 2>> $OUT/${TAG}_bench_n$N.err; echo "ref rc=$?"; cat $OUT/${TAG}_ref_n$N.json
