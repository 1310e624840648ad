#!/bin/bash
# second box visit: fixed headless tests, L2-policy A/B, full GPU suite
set -u
O=gpurun_out; mkdir -p $O
echo "== headless tests"; timeout 300 python -m pytest tests/test_gpu_headless.py -m gpu -q 2>&1 | tail -15 | tee $O/pytest_headless.txt
echo "== A/B"; scripts/ab_variants.sh "hl" base v4 v7 v12 v13 2>&1 | tee $O/ab_l2_policy.txt
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 | tee $O/pytest_gpu.txt
