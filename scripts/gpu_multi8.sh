#!/bin/bash
# 8-GPU check exactly as the driver launches it: reference arm (rank 0 only), then the b200 arm
N=${1:-8}; OUT=gpurun_out/${2:-m08}
mkdir -p $OUT
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_n${N}_reference.json 2> $OUT/bench_n${N}_reference.err; echo "rc=$?"; cut -c1-300 $OUT/bench_n${N}_reference.json
echo "== b200 arm N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_fused.json 2> $OUT/bench_n${N}_fused.err; echo "rc=$?"; cut -c1-1200 $OUT/bench_n${N}_fused.json; grep -v "OMP_NUM\|^\*\*\*\|^$" $OUT/bench_n${N}_fused.err | tail -5
