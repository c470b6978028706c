#!/bin/bash
# one ncu --set full capture of a named kernel from a short bench run; details page as text
OUT=gpurun_out/${1:-n1}; K=${2:-k_iekf_next}; SKIP=${3:-4}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/$K -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $OUT/b_$K.log 2>&1
ncu -i $OUT/$K.ncu-rep --page details > $OUT/${K}_details.txt 2>&1
grep -E "Duration|Registers Per|Executed Ipc|Issue Slots Busy|No Instruction|Stall|Warp Cycles Per Issued|Instructions Executed|Theoretical Occ|Achieved Occ" $OUT/${K}_details.txt | head -40
ncu -i $OUT/$K.ncu-rep --page raw --csv 2>/dev/null > $OUT/${K}_raw.csv
python - <<PY
import sys,csv
rows=list(csv.reader(open("$OUT/${K}_raw.csv")))
if len(rows)>=3:
    hdr=rows[0]; val=rows[-1]
    for h,v in zip(hdr,val):
        if "smsp__average_warp" in h and "issue_stalled" in h and "_per_warp_active" in h and "not_issued" not in h: print(h.split("issue_stalled_")[1][:40], v)
        if h in ("smsp__inst_executed.sum","sm__inst_executed.sum","smsp__warps_launched.sum"): print(h,v)
PY
