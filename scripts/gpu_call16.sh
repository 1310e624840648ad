#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/measure_configs.py frame c1 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-120s %8.2f Gcell/s  %.5f ms  %s' % (d['config'][:120], d['gcell_per_s'], d['ms_per_step'], d['note'][:90]))" | tee $OUT/r02_frame_loop.txt
