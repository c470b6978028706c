#!/bin/bash
# the driver's multi-GPU launch of the b200 arm, default steps
N=${1:-2}; OUT=gpurun_out/${2:-mq}
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 50 --warmup 3 > $OUT/bench_n${N}.json 2> $OUT/bench_n${N}.err; echo "rc=$?"; cut -c1-400 $OUT/bench_n${N}.json; grep -v "OMP_NUM\|^\*\*\*\|^$" $OUT/bench_n${N}.err | tail -4
