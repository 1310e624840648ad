#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash scripts/profile.sh r02 --no-parity-check --no-other-configs > $OUT/r02_profile.log 2>&1
ncu -i $OUT/r02_march.ncu-rep --page source --csv > $OUT/r02_march_source.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/r02_march_raw.csv > $OUT/r02_k_march_tma_lut.txt 2>&1; tail -40 $OUT/r02_k_march_tma_lut.txt
timeout 300 python scripts/measure_configs.py c1 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-90s %8.2f Gcell/s  %.5f ms' % (d['config'][:90], d['gcell_per_s'], d['ms_per_step']))"
