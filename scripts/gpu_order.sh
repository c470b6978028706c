#!/bin/bash
# the single-launch sweep ordering: host-clock A/B of the variants, per-kernel times (ncu), then two default bench runs
OUT=gpurun_out/${1:-o1}
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 200 python scripts/order_time.py 100000 200 2>&1 | tee $OUT/order_time.txt | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/order_launches.csv python scripts/order_time.py 100000 2 > $OUT/ncu.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/order_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items():
    print("%-62s n=%3d avg %.1f us  min %.1f max %.1f"%(k,len(v),sum(v)/len(v)/1000,min(v)/1000,max(v)/1000))
PY
for v in a b; do
timeout 300 python bench.py --no-cpu-baseline --steps 40 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; tail -2 $OUT/bench_$v.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$v.json")); print("run $v value %.1fM e2e %.1fM ms/step %.3f e2e ms %.3f | %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"],d["config"]["sweep_order"][:40])); print(d.get("ms_per_step_stats"))
except Exception as e: print("no bench line", e)
PY
done
