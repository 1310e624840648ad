#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the dominant kernel.
# Usage: scripts/profile.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_march -s 4 -c 2 -f -o $OUT/${TAG}_march \
    python bench.py --steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_march.ncu-rep --page raw --csv > $OUT/${TAG}_march_raw.csv 2>/dev/null
ls -la $OUT | tail -8
