#!/bin/bash
# compute-sanitizer over every kernel of the hot path at small shapes (64^3, 320x240); logs -> $1 (default gpurun_out/san)
# memcheck + racecheck + initcheck + synccheck on the ordinary kernels; memcheck + racecheck on the persistent ICP
# kernel with its poll bound raised to 120 s (it talks to the host while it runs).
OUT=${1:-gpurun_out/san}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
for tool in memcheck racecheck initcheck synccheck; do
  KFB_SAN_NO_PERSISTENT=1 timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python tools/sanitize_driver.py > "$OUT/${tool}_direct.log" 2>&1
  echo "$tool direct rc=$?" | tee -a "$OUT/summary.txt"
done
for tool in memcheck racecheck; do
  KFB_ICP_TIMEOUT_NS=120000000000 timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
    python tools/sanitize_driver.py > "$OUT/${tool}_persistent.log" 2>&1
  echo "$tool persistent rc=$?" | tee -a "$OUT/summary.txt"
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|ok:" "$OUT"/*.log | tee -a "$OUT/summary.txt"
