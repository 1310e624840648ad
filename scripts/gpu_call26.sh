#!/bin/bash
# Where do the ~5 us per step of the sharded 1024-row launch go?  Timing-only variant (-DSBX_PEER_DEBUG) with parts of the
# peer protocol switched off (results are wrong in those runs; only the times are read).
OUT=gpurun_out; mkdir -p $OUT
D=soap_bubble_simulation_b200/csrc
cp $D/libsbx.so /tmp/base.so; cp $D/ab/libsbx_peerdbg.so $D/libsbx.so
ROWS=2048
run() { local label=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --scaling strong --rows $ROWS --steps 400 --warmup 20 --no-parity-check --no-other-configs --no-cpu-baseline --e2e-steps 0 2>/dev/null \
    | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-40s rows=%s  %9.2f Gcell/s  %.4f ms' % ('$label', '$ROWS', d['value'], d['ms_per_step']))"
}
{ run "full protocol" SBX_PEER_DEBUG_BITS=0
  run "no wait" SBX_PEER_DEBUG_BITS=1
  run "no push" SBX_PEER_DEBUG_BITS=2
  run "no system fence" SBX_PEER_DEBUG_BITS=8
  run "no wait, no signal" SBX_PEER_DEBUG_BITS=5
  run "nothing (wait, push, signal off)" SBX_PEER_DEBUG_BITS=7
  run "full protocol, no PDL" SBX_PEER_DEBUG_BITS=0 SBX_NO_PDL=1
  run "nothing, no PDL" SBX_PEER_DEBUG_BITS=7 SBX_NO_PDL=1
  run "full protocol again" SBX_PEER_DEBUG_BITS=0
  echo "single GPU, same rows per GPU:"
  timeout 200 python scripts/measure_configs.py grid:1024:16384 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-60s %8.2f Gcell/s  %.4f ms' % (d['config'][:60], d['gcell_per_s'], d['ms_per_step']))"
} > $OUT/r02_peer_protocol_cost.txt 2>&1
cp /tmp/base.so $D/libsbx.so
cat $OUT/r02_peer_protocol_cost.txt
