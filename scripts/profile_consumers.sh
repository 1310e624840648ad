#!/bin/bash
# Runs on the GPU box (under gpurun): one full ncu capture each of the ring consumers (shade-only, caustic map).
set -u
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:k_shade_tma -s 3 -c 1 -f -o $OUT/r01_shade_tma \
    python scripts/measure_configs.py c5 > $OUT/r01_shade_tma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_caustics_tma -s 3 -c 1 -f -o $OUT/r01_caustics_tma \
    python scripts/measure_configs.py f2 > $OUT/r01_caustics_tma.log 2>&1
for k in shade_tma caustics_tma; do
  ncu -i $OUT/r01_$k.ncu-rep --page raw --csv > $OUT/r01_${k}_raw.csv 2>/dev/null
done
ls -la $OUT | tail -8
