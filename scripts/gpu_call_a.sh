#!/bin/bash
# One GPU-box visit: new-row tests, store-hint A/B, bench (both arms), ncu refresh, then the full GPU suite.
set -u
O=gpurun_out; mkdir -p $O
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== headless tests"; timeout 300 python -m pytest tests/test_gpu_headless.py tests/test_reference_golden_images.py -m gpu -q 2>&1 | tail -15
echo "== A/B"; scripts/ab_variants.sh "hl" base v1 v2 v3 2>&1 | tee $O/ab_store_hint.txt
echo "== bench"; timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
echo "== ncu"; timeout 400 scripts/profile.sh r01c > $O/profile.log 2>&1; tail -3 $O/profile.log
echo "== full gpu suite"; timeout 500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 | tee $O/pytest_gpu.txt
