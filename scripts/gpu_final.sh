#!/bin/bash
# end-of-round capture on one GPU: tests, smoke, both bench arms, ncu launch list, ncu full captures (100k sweep and a
# 12.5k-keypoint sweep = an 8-GPU shard), host-loop A/B, config-5-sized run.  ncu runs use the host-driven loop (a replayed
# kernel cannot take part in the persistent ESIKF block's hand-over; the library detects serialised kernels itself).
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cut -c1-300 $OUT/bench_reference.json
echo "== bench b200"; timeout 900 python bench.py > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "rc=$?"; cut -c1-1500 $OUT/bench_b200.json; tail -3 $OUT/bench_b200.err
echo "== bench b200, host-driven loop (round-1 form of updateIEKF)"; SRL_DEVICE_LOOP=0 timeout 600 python bench.py --no-cpu-baseline --steps 40 > $OUT/bench_hostloop.json 2> $OUT/bench_hostloop.err; echo "rc=$?"
echo "== bench b200, 12.5k-keypoint sweep (the shard of an 8-GPU run)"; timeout 600 python bench.py --no-cpu-baseline --steps 40 --points 12500 > $OUT/bench_12k5.json 2> $OUT/bench_12k5.err; echo "rc=$?"
echo "== ncu launch list"
SRL_DEVICE_LOOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full captures"
export SRL_DEVICE_LOOP=0
for K in k1_scan k1_fit; do for P in 100000 12500; do
T=${K}_$P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o /tmp/$T -f python bench.py --no-cpu-baseline --steps 2 --warmup 1 --points $P > $OUT/b_$T.log 2>&1
ncu -i /tmp/$T.ncu-rep --page details > $OUT/${T}_details.txt 2>&1
ncu -i /tmp/$T.ncu-rep --page raw --csv > $OUT/${T}_raw.csv 2>/dev/null
done; done
unset SRL_DEVICE_LOOP
echo "== config-5-sized run on one GPU: 500k-pt spinning sweep, ~50M-pt map, 5 passes"
timeout 1500 python bench.py --points 500000 --map-extent 1340 --passes 5 --pattern spinning --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_1gpu.json 2> $OUT/bench_cfg5_1gpu.err; echo "rc=$?"; cut -c1-600 $OUT/bench_cfg5_1gpu.json; tail -3 $OUT/bench_cfg5_1gpu.err
du -sh $OUT
