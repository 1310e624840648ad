#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/ab_env.sh SBX_RING_STAGES "3 2 0" "hl c3 c4 l2ab" > $OUT/r02_ab_ring_stages.txt 2>&1; cat $OUT/r02_ab_ring_stages.txt
