#!/bin/bash
# Same-box A/B of several library builds (box-to-box variance is several per cent, so only same-box runs compare):
#   scripts/ab_variants.sh "<measure_configs.py selectors>" base v1 v2 ...
# "base" is csrc/libsbx.so, vN is csrc/ab/libsbx_vN.so (built beforehand, e.g. with -DSBX_STORE_HINT=N).
D=soap_bubble_simulation_b200/csrc
SEL=$1; shift
cp $D/libsbx.so /tmp/base.so
for rep in 1 2; do
  for v in "$@"; do
    if [ $v = base ]; then cp /tmp/base.so $D/libsbx.so; else cp $D/ab/libsbx_$v.so $D/libsbx.so; fi
    echo "== $v (rep $rep)"
    timeout 300 python scripts/measure_configs.py $SEL 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-70s %8.2f Gcell/s  %.4f ms' % (d['config'][:70], d['gcell_per_s'], d['ms_per_step']))"
  done
done
cp /tmp/base.so $D/libsbx.so
