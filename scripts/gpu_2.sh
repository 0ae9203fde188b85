#!/bin/bash
# Two-rank run: weak-scaling bench of the inference pass and of the training step (one NCCL all-reduce of the flat
# 202.5 MB gradient buffer per step).
TAG=${1:-r1z2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench_train.py --gpus 2 --steps 10 --warmup 3 --cpu-steps 0 > $OUT/bench_train_2gpu.json 2> $OUT/bench_train_2gpu.err
echo "train exit $?"; cat $OUT/bench_train_2gpu.json; tail -3 $OUT/bench_train_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
echo "bench exit $?"; cat $OUT/bench_2gpu.json | cut -c1-700; tail -3 $OUT/bench_2gpu.err
