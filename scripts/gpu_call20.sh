#!/bin/bash
OUT=gpurun_out
{ for br in 0 24 37 48 56 64; do for keep in 1 0; do
  echo "== SBX_BAND_ROWS=$br SBX_L2_KEEP=$keep"
  SBX_BAND_ROWS=$br SBX_L2_KEEP=$keep timeout 300 python scripts/measure_configs.py exact_hl 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
done; done
echo "== SBX_NO_PDL=1"; SBX_NO_PDL=1 python scripts/measure_configs.py exact_hl 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
} > $OUT/r02_ab_band_l2.txt 2>&1
cat $OUT/r02_ab_band_l2.txt
