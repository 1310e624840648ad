#!/bin/bash
# Strong scaling at 1024 rows per GPU, two ranks: does a short edge band pay in a single-wave launch?
OUT=gpurun_out; mkdir -p $OUT
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --scaling strong --rows $ROWS --steps 400 --warmup 20 --no-parity-check --no-other-configs --no-cpu-baseline --e2e-steps 1 --halo ${HALO:-peer} 2>/dev/null \
    | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-40s rows=%s  %9.2f Gcell/s  %.4f ms' % ('$label', '$ROWS', d['value'], d['ms_per_step']))"
}
{ for ROWS in 2048 4096; do
  run "default" SBX_X=0
  run "edge8 always" SBX_EDGE_MIN_WAVES=0
  run "edge4 always" SBX_EDGE_MIN_WAVES=0 SBX_EDGE_ROWS=4
  run "edge12 always" SBX_EDGE_MIN_WAVES=0 SBX_EDGE_ROWS=12
  run "no edge" SBX_EDGE_ROWS=0
  HALO=nccl run "nccl halo" SBX_X=0; HALO=peer
done; } > $OUT/r02_ab_edge_single_wave.txt 2>&1
cat $OUT/r02_ab_edge_single_wave.txt
