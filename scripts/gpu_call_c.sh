#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== default"; timeout 300 python scripts/diag_c3.py 6 2>&1 | tee $O/diag_default.txt
echo "== SBX_NO_PDL=1"; SBX_NO_PDL=1 timeout 200 python scripts/diag_c3.py 4 2>&1 | tee $O/diag_nopdl.txt
echo "== SBX_BAND_ROWS=24"; SBX_BAND_ROWS=24 timeout 200 python scripts/diag_c3.py 4 2>&1 | tee $O/diag_band24.txt
echo "== the failing test, 3 times"; for i in 1 2 3; do timeout 200 python -m pytest "tests/test_gpu_fullsize.py::test_full_size_fused_step_matches_oracle" -q -m gpu 2>&1 | tail -3; done
