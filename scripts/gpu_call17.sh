#!/bin/bash
OUT=gpurun_out
{ bash scripts/ab_variants.sh hl base st2 st4; echo "--- st2 with carveout 58 (132 KB), st4 with carveout 100 (228 KB)";
  D=soap_bubble_simulation_b200/csrc; cp $D/libsbx.so /tmp/base.so
  cp $D/ab/libsbx_st2.so $D/libsbx.so; SBX_SMEM_CARVEOUT=58 python scripts/measure_configs.py hl 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   st2 c58 %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
  cp $D/ab/libsbx_st4.so $D/libsbx.so; SBX_SMEM_CARVEOUT=100 python scripts/measure_configs.py hl 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   st4 c100 %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
  cp /tmp/base.so $D/libsbx.so; } > $OUT/r02_ab_stages.txt 2>&1
cat $OUT/r02_ab_stages.txt
