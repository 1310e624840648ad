#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/measure_configs.py hl c3 c4 l2ab wgsl 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-70s %8.2f Gcell/s  %.4f ms  %.1f%%' % (d['config'][:70], d['gcell_per_s'], d['ms_per_step'], 100*d['frac_of_measured_hbm']))" | tee $OUT/r02_configs_final.txt
