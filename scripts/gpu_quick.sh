#!/bin/bash
# quick check of a kernel change: the split / randomized parity tests, one bench run, per-kernel launch times
OUT=gpurun_out/${1:-q1}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -k "split or hands_ambiguous or randomized or pass_matches" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $OUT/bench$i.json 2> $OUT/bench$i.err
python - <<PY
import json
d=json.load(open("$OUT/bench$i.json")); print("value %.1fM e2e %.1fM ms/step %.3f k1 %.4f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["roofline"]["k1_avg_ms"]))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $OUT/b.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items():
    if "k1_" in k: print("%-62s n=%3d avg %.1f us"%(k,len(v),sum(v)/len(v)/1000))
PY
