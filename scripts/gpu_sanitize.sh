#!/bin/bash
# compute-sanitizer passes: the small diagnostic (map build, insert, scan-matching passes, ESIKF update) under racecheck /
# synccheck / initcheck, and the tests of the newer entry points under memcheck
TAG=${1:-san1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for tool in racecheck synccheck initcheck memcheck; do
  echo "== $tool (gpu_diag small)"; timeout 700 compute-sanitizer --tool $tool python scripts/gpu_diag.py small > $OUT/$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/$tool.log | head -3
done
echo "== memcheck (pytest: split pass, undistortion, eviction, grid sampling)"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -x -k "split or undistortion or remove_points or grid_sampling_matches" > $OUT/memcheck_pytest.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck_pytest.log | tail -3
