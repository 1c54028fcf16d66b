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
  limit=40
  [ "$tool" = racecheck ] && extra="--racecheck-report all" && limit=1000000
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit $limit $extra \
      python tools/sanitize_target.py > $out 2>&1
  code=$?
  echo "== $tool exit $code" | tee -a $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_target ok|Error|hazard" $out | tail -15
  [ $code -ne 0 ] && rc=1
done
# racecheck: one line per (kind, kernel, shared address) instead of megabytes of repeated backtraces
python - <<'PY' > gpurun_out/sanitizer_racecheck_summary.txt
import collections, re
kinds = collections.Counter()
cur = None
for line in open("gpurun_out/sanitizer_racecheck.txt", errors="replace"):
    m = re.search(r"(Error|Warning): (.*?) at __shared__ (0x[0-9a-f]+) in block", line)
    if m:
        cur = [m.group(2), m.group(3), None]
        continue
    m = re.search(r"(Write|Read) Thread .*? at (?:void )?([\w:]+)", line)
    if m and cur is not None and cur[2] is None:
        cur[2] = m.group(2)
        kinds[tuple(cur)] += 1
for (kind, addr, kern), n in sorted(kinds.items(), key=lambda kv: (kv[0][2], kv[0][0], int(kv[0][1], 16))):
    print("%6d  %-60s %-10s %s" % (n, kind, addr, kern))
print("total", sum(kinds.values()))
PY
mv gpurun_out/sanitizer_racecheck.txt gpurun_out/sanitizer_racecheck_full.txt
head -c 20000 gpurun_out/sanitizer_racecheck_full.txt > gpurun_out/sanitizer_racecheck.txt
tail -5 gpurun_out/sanitizer_racecheck_full.txt >> gpurun_out/sanitizer_racecheck.txt
rm -f gpurun_out/sanitizer_racecheck_full.txt
exit $rc
