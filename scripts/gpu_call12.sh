#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02c_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02c_pytest_gpu.txt; tail -6 $OUT/r02c_pytest_gpu.txt
rm -f $OUT/sanitize_summary.txt
for tool in memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --launch-timeout 120 --error-exitcode 9 \
      python scripts/sanitize_target.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/sanitize_summary.txt
  grep -E "ERROR SUMMARY|SANITIZE_TARGET_DONE|========= (Invalid|Race|Barrier|Error)" $OUT/sanitize_$tool.log | head -8 | tee -a $OUT/sanitize_summary.txt
done
timeout 600 python bench.py --workload foam --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02_bench_foam_n1.json 2> $OUT/r02_bench_foam_n1.err; echo "foam rc=$?"
python -c "
import json; d=json.load(open('$OUT/r02_bench_foam_n1.json')); print('foam', d['value'], d['ms_per_step'], d['e2e'], d.get('parity_check',{}).get('ok'))"
