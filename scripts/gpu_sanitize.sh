#!/bin/bash
# compute-sanitizer passes over the round-2 code: the small diagnostic (map build, insert, passes, device-resident ESIKF
# loop) under racecheck / synccheck / initcheck / memcheck, and the tests of the newer entry points under memcheck.
# The library reports whether the device-resident loop was usable under the tool (a tool that serialises kernels makes it
# fall back to the host-driven loop).
TAG=${1:-san2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for tool in racecheck synccheck initcheck memcheck; do
  echo "== $tool (gpu_diag small)"; timeout 700 compute-sanitizer --tool $tool python scripts/gpu_diag.py small > $OUT/$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|device loop" $OUT/$tool.log | head -4
done
echo "== memcheck (pytest: device loop, colour map, split pass, cap, grid sampling)"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -x -k "device_resident or color_map or split or cap or grid_sampling_matches or iekf_matches" > $OUT/memcheck_pytest.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck_pytest.log | tail -3
echo "== racecheck (pytest: device loop + colour map)"
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests -q -m gpu -x -k "device_resident or color_map" > $OUT/racecheck_pytest.log 2>&1; echo "rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" $OUT/racecheck_pytest.log | tail -3
