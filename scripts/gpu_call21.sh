#!/bin/bash
OUT=gpurun_out
{ for br in 0 12 16 20 24 28 32 37 42; do
  echo "== SBX_BAND_ROWS=$br"
  SBX_BAND_ROWS=$br timeout 300 python scripts/measure_configs.py exact_hl 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
done
for st in 3; do for br in 24 37; do
  echo "== SBX_RING_STAGES=$st SBX_BAND_ROWS=$br"
  SBX_RING_STAGES=$st SBX_BAND_ROWS=$br timeout 300 python scripts/measure_configs.py exact_hl 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
done; done
} > $OUT/r02_ab_band_rows.txt 2>&1
cat $OUT/r02_ab_band_rows.txt
