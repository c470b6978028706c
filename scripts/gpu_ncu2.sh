#!/bin/bash
# ncu --set full captures of k1_scan and k1_fit on the 100k sweep, and of both on a 12.5k-keypoint sweep (an 8-GPU shard).
# Host-driven loop (SRL_DEVICE_LOOP=0): under a profiler that replays a kernel the persistent ESIKF block cannot take part.
OUT=gpurun_out/${1:-n2}
mkdir -p $OUT
export SRL_DEVICE_LOOP=0
for K in k1_scan k1_fit; do
for P in 100000 12500; do
T=${K}_$P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o /tmp/$T -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 --points $P > $OUT/b_$T.log 2>&1
ncu -i /tmp/$T.ncu-rep --page details > $OUT/${T}_details.txt 2>&1
ncu -i /tmp/$T.ncu-rep --page raw --csv > $OUT/${T}_raw.csv 2>/dev/null
if [ $P = 100000 ]; then ncu -i /tmp/$T.ncu-rep --page source --csv > $OUT/${T}_source.csv 2>/dev/null; fi
done
done
for f in $OUT/*_details.txt; do echo "== $f"; grep -E "^\s+Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|Avg. Active Threads|Executed Instructions  |DRAM Throughput" $f | head -12; done
python - <<PY
import csv
for K in ("k1_scan_100000","k1_fit_100000","k1_scan_12500","k1_fit_12500"):
    rows=list(csv.reader(open("$OUT/%s_raw.csv"%K)))
    hdr,val=rows[0],rows[-1]
    d=dict(zip(hdr,val))
    print(K, {k:d[k] for k in d if k in ("dram__bytes_read.sum","dram__bytes_write.sum","smsp__inst_executed.sum","gpu__time_duration.sum","sm__warps_active.avg.pct_of_peak_sustained_active")})
    st={}
    for k,v in d.items():
        if "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
            try: st[k.split("issue_stalled_")[1].replace("_per_warp_active.pct","")]=float(v)
            except: pass
    print("  stalls%:", sorted(st.items(), key=lambda x:-x[1])[:8])
PY
du -sh $OUT
