#!/bin/bash
# A/B of two library builds on the same box (box-to-box variance is several per cent, so only same-box pairs compare):
#   scripts/ab_run.sh <measure_configs.py arguments...>
# compares csrc/libsbx.so with csrc/ab/libsbx_old.so, which you build from the old source first, e.g.
#   nvcc $ARCH ... -c old/sbx_march_tma.cu -o ab/old.o && nvcc $ARCH -shared -o ab/libsbx_old.so <other objects> ab/old.o -ldl
D=soap_bubble_simulation_b200/csrc
cp $D/libsbx.so /tmp/new.so
for rep in 1 2; do
  for v in new old; do
    if [ $v = new ]; then cp /tmp/new.so $D/libsbx.so; else cp $D/ab/libsbx_old.so $D/libsbx.so; fi
    echo "== $v (rep $rep)"
    timeout 600 python scripts/measure_configs.py "$@" 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-70s %8.2f Gcell/s' % (d['config'][:70], d['gcell_per_s']))"
  done
done
cp /tmp/new.so $D/libsbx.so
