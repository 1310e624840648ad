#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== minmax tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "minmax or critical or api" 2>&1 | tail -6
echo "== frame loop"; timeout 300 python scripts/measure_configs.py frame 2>&1 | grep "^{\|Error\|error" | tee $O/configs_frame.txt
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
