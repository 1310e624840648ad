#!/bin/bash
# Mid-size band-height rule: same-box check of the new automatic choice, then the GPU suite, capture and bench.
OUT=gpurun_out; mkdir -p $OUT
{ for g in grid:4096:16384 grid:2048:16384 grid:1024:16384 grid:2048:4096 hl c2 c4; do
  echo "== $g auto"
  timeout 300 python scripts/measure_configs.py $g 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
done; } > $OUT/r02_band_rows_auto_after.txt 2>&1
cat $OUT/r02_band_rows_auto_after.txt
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py > $OUT/r02_pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_final.txt
bash scripts/gpu_call15.sh
