#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== A/B L2 keep"; scripts/ab_env.sh SBX_L2_KEEP "0 1" "hl l2ab c4" 2>&1 | tee $O/ab_l2_keep_env.txt
echo "== fullsize tests (default policy)"; timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
