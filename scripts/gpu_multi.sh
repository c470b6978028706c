#!/bin/bash
# multi-GPU bench exactly as the driver launches it
N=${1:-2}; TAG=${2:-m02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== b200 arm N=$N fused peer-memory exchange"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_fused.json 2> $OUT/bench_n${N}_fused.err; echo "rc=$?"; cat $OUT/bench_n${N}_fused.json; grep -v "OMP_NUM\|^\*\*\*\|^$" $OUT/bench_n${N}_fused.err | tail -5
echo "== b200 arm N=$N NCCL all-reduce from Python (baseline)"
SRL_DIST_NATIVE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_nccl.json 2> $OUT/bench_n${N}_nccl.err; echo "rc=$?"; cat $OUT/bench_n${N}_nccl.json
echo "== 2-rank pytest (fused exchange across real GPUs)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k fused_peer > $OUT/pytest_ipc.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_ipc.log
