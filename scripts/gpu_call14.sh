#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scripts/ab_env.sh SBX_PK "1 2" "hl c3 c4 tfe" > $OUT/r02_ab_pk.txt 2>&1; cat $OUT/r02_ab_pk.txt
