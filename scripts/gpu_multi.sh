#!/bin/bash
# multi-GPU bench exactly as the driver launches it
N=${1:-2}; TAG=${2:-m01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
echo "== reference arm under torchrun N=$N (rank 0 only)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "rc=$?"; tail -c 600 $OUT/bench_ref_n$N.json
echo "== b200 arm N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; cat $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
echo "== b200 arm N=1 (same box)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; cat $OUT/bench_n1.json
