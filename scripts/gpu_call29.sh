#!/bin/bash
# Reversed march of the band under an up neighbour (few-wave launches): parity at two ranks, then the strong-scaling shapes.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "peer and not wgsl" > $OUT/r02_pytest_multi_rev_n2.txt 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02_pytest_multi_rev_n2.txt
run() { local label=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --scaling strong --rows $ROWS --steps 400 --warmup 20 --no-other-configs --no-cpu-baseline --e2e-steps 0 2>/dev/null \
    | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-40s rows=%s  %9.2f Gcell/s  %.4f ms  parity %s' % ('$label', '$ROWS', d['value'], d['ms_per_step'], d.get('parity_check', {}).get('ok')))"
}
{ ROWS=2048; run "reversed top band" SBX_X=0
  ROWS=4096; run "reversed top band" SBX_X=0
} > $OUT/r02_rev_strong_n2.txt 2>&1
cat $OUT/r02_rev_strong_n2.txt
