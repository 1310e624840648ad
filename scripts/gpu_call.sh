#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== new TFE test"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fronts" 2>&1 | tail -6
echo "== tfe timing"; timeout 300 python scripts/measure_configs.py tfe hl 2>&1 | grep "^{" | tee $O/configs_tfe.txt
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
