#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
