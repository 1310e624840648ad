#!/bin/bash
OUT=gpurun_out
{ bash scripts/ab_env.sh SBX_SHADE_CARVEOUT "100 86 72" c5; bash scripts/ab_env.sh SBX_WGSL_CARVEOUT "100 86 72" wgsl; bash scripts/ab_env.sh SBX_SMEM_CARVEOUT "100 72" "l2ab c4"; } > $OUT/r02_ab_carveout2.txt 2>&1
cat $OUT/r02_ab_carveout2.txt
