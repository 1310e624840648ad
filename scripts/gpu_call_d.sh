#!/bin/bash
# ring-race fix: determinism regression + full suite; then perf A/B of the fixed base, L2 hint and TEX-pipelined variants
set -u
O=gpurun_out; mkdir -p $O
D=soap_bubble_simulation_b200/csrc
echo "== determinism + fullsize"; timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 | tee $O/pytest_fullsize.txt
echo "== A/B"; scripts/ab_variants.sh "hl" base v4 p4 p3 2>&1 | tee $O/ab_fix_pipe.txt
cp $D/libsbx.so /tmp/keep.so
for v in v4 p4 p3; do
  echo "== correctness of $v"; cp $D/ab/libsbx_$v.so $D/libsbx.so
  timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "fused_step_matches_oracle or (deterministic and cpu_f32) or tma_and_direct" 2>&1 | tail -4
done
cp /tmp/keep.so $D/libsbx.so
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
