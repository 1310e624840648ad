#!/bin/bash
# Same-box A/B of one environment knob of libsbx:  scripts/ab_env.sh SBX_L2_KEEP "0 1" "<measure_configs selectors>"
VAR=$1; VALS=$2; SEL=$3
for rep in 1 2; do
  for v in $VALS; do
    echo "== $VAR=$v (rep $rep)"
    env $VAR=$v timeout 400 python scripts/measure_configs.py $SEL 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-70s %8.2f Gcell/s  %.4f ms' % (d['config'][:70], d['gcell_per_s'], d['ms_per_step']))"
  done
done
