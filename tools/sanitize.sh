#!/bin/bash
# compute-sanitizer over the hot-path kernels at small shapes (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# Summaries land in gpurun_out/sanitizer_<tool>.txt; copy them to profiles/ (named per round) when clean.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RVB_SANITIZE_FRAMES=${RVB_SANITIZE_FRAMES:-160}
rc=0
for tool in memcheck racecheck synccheck initcheck; do
  out=gpurun_out/sanitizer_${tool}.txt
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 7 --print-limit 40 \
      python tools/sanitize_target.py > $out 2>&1
  code=$?
  echo "== $tool exit $code" | tee -a $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_target ok|Error|hazard" $out | tail -15
  [ $code -ne 0 ] && rc=1
done
exit $rc
