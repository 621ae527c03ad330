#!/bin/bash
# compute-sanitizer passes over tests/bringup/sanitize_run.py (every C-ABI entry point on small volumes, oracle-checked).
# usage (on a GPU box): tools/sanitize.sh [tool ...]     default: memcheck racecheck
# Output: gpurun_out/sanitize_<tool>.log; the last lines of each log are echoed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools=${@:-memcheck racecheck}
for t in $tools; do
  extra=""
  [ "$t" = memcheck ] && extra="--leak-check no"
  small=${SANITIZE_SMALL:-0}; [ "$t" = racecheck ] && small=1
  SANITIZE_SMALL=$small timeout ${SANITIZE_TIMEOUT:-200} compute-sanitizer --tool $t $extra --error-exitcode 9 --print-limit 30 \
    python tests/bringup/sanitize_run.py > gpurun_out/sanitize_$t.log 2>&1
  echo "== $t rc=$? =="
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok|Error|hazard" gpurun_out/sanitize_$t.log | sort | uniq -c | sort -rn | head -12
done
