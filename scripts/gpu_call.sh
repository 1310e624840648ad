#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/sanitize_summary.txt
for tool in memcheck synccheck; do
  timeout 200 compute-sanitizer --tool $tool --launch-timeout 120 --error-exitcode 9 \
      python scripts/sanitize_target.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/sanitize_summary.txt
  grep -E "ERROR SUMMARY|SANITIZE_TARGET_DONE|========= (Invalid|Barrier|Error)" $OUT/sanitize_$tool.log | head -8 | tee -a $OUT/sanitize_summary.txt
done
