#!/bin/bash
# compute-sanitizer passes over every kernel family (run under gpurun); summaries land in gpurun_out/.
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --launch-timeout 120 --error-exitcode 9 \
      python scripts/sanitize_target.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/sanitize_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_TARGET_DONE|========= (Invalid|Race|Barrier|Error)" $OUT/sanitize_$tool.log | head -8 | tee -a $OUT/sanitize_summary.txt
done
