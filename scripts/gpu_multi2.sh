#!/bin/bash
# N-GPU check: multi-rank tests, bench at N ranks (device loop, with the cfg5 block), host-loop comparison without cfg5
N=${2:-2}
OUT=gpurun_out/${1:-m2}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x -k "exchange_ranks" > $OUT/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 $OUT/pytest_dist.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 --cfg5-steps 6 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 $OUT/bench_n$N.err
SRL_DEVICE_LOOP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 3 --no-cfg5 > $OUT/bench_n${N}_hostloop.json 2> $OUT/bench_n${N}_hostloop.err; echo "bench hostloop rc=$?"
python - <<PY
import json
for f in ("$OUT/bench_n$N.json","$OUT/bench_n${N}_hostloop.json"):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f k1 %.4f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["k1_avg_ms"]), d.get("pose_check"))
        if "cfg5" in d: c=d["cfg5"]; print("  cfg5 value %.1fM e2e %.1fM ms/step %.3f frac %.3f gen %s s"%(c["value"]/1e6,c["e2e"]["value"]/1e6,c["ms_per_step"],c["roofline"]["frac"],c["map_gen_s"]))
    except Exception as e: print("no bench line", f, e)
PY
