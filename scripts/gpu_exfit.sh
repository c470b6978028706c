#!/bin/bash
# 2-GPU validation of the exchange run by k1_fit (SRL_EXCHANGE_IN_FIT=1): the 2-rank parity test, then the bench A/B
OUT=gpurun_out/${1:-exfit}
mkdir -p $OUT
SRL_EXCHANGE_IN_FIT=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k fused_peer > $OUT/pytest_ipc.log 2>&1; echo "pytest(exchange in fit) rc=$?"; tail -3 $OUT/pytest_ipc.log
for m in 1 0; do
SRL_EXCHANGE_IN_FIT=$m timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$m bench.py --gpus 2 --steps 30 --warmup 3 > $OUT/bench_n2_m$m.json 2> $OUT/bench_n2_m$m.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n2_m$m.json")); print("exchange_in_fit=$m value %.1fM e2e %.1fM ms/step %.3f k1 %.4f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"]))
except Exception as e: print("failed", e)
PY
done
