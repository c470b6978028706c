#!/bin/bash
# end-of-round capture: tests, smoke, both bench arms, ncu launch list + one full capture, config-5-sized run
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cut -c1-400 $OUT/bench_reference.json
echo "== bench b200"; timeout 900 python bench.py > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "rc=$?"; cat $OUT/bench_b200.json; tail -3 $OUT/bench_b200.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full k1_scan + k1_fit (the 3 passes of one sweep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_scan|k1_fit" -s 6 -c 6 -f -o $OUT/k1_split_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_split.log 2>&1; echo "rc=$?"
echo "== A/B: previous form of the pass (k1_fast)"
timeout 300 python bench.py --no-cpu-baseline --steps 20 --k1-variant 1 > $OUT/bench_k1fast.json 2> $OUT/bench_k1fast.err; echo "rc=$?"
echo "== config-5-sized run on one GPU: 500k-pt spinning sweep, ~50M-pt map, 5 passes"
timeout 1500 python bench.py --points 500000 --map-extent 1340 --passes 5 --pattern spinning --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_1gpu.json 2> $OUT/bench_cfg5_1gpu.err; echo "rc=$?"; cat $OUT/bench_cfg5_1gpu.json; tail -3 $OUT/bench_cfg5_1gpu.err
