#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== A/B kernel with/without by-product code (tracking off)"; scripts/ab_variants.sh "hl" base nomm 2>&1 | tee $O/ab_mm_track_codegen.txt
echo "== frame loop, tracking off / on"; scripts/ab_env.sh SBX_TRACK_MINMAX "0 1" "frame" 2>&1 | tee $O/ab_mm_track_frame.txt
