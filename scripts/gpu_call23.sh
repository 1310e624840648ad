#!/bin/bash
OUT=gpurun_out
{ for g in grid:4096:16384 grid:2048:16384 grid:1024:16384 grid:2048:4096; do for br in 0 14 20 24 28 32; do
  echo "== $g SBX_BAND_ROWS=$br"
  SBX_BAND_ROWS=$br timeout 300 python scripts/measure_configs.py $g 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
done; done; } > $OUT/r02_ab_band_rows_mid.txt 2>&1
cat $OUT/r02_ab_band_rows_mid.txt
