#!/bin/bash
# N-GPU check of the short / first-dispatched boundary bands (peer push) + strong scaling line + resident frame timing
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "peer or host" > $OUT/r02c_pytest_multi_n$N.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02c_pytest_multi_n$N.txt
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $OUT/r02c_bench_${name}_n$N.json 2> $OUT/r02c_bench_${name}_n$N.err
  echo "bench $name rc=$?"; python -c "
import json; d=json.load(open('$OUT/r02c_bench_${name}_n$N.json')); print('$name', d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['ms_per_step'], d.get('parity_check',{}).get('ok'))"; }
run strong --scaling strong --steps 200 --warmup 20 --no-other-configs
SBX_EDGE_ROWS=0 run strong_noedge --scaling strong --steps 200 --warmup 20 --no-other-configs --e2e-steps 0
run strong1k --scaling strong --rows $((1024*N)) --steps 400 --warmup 40 --no-other-configs --e2e-steps 0
SBX_EDGE_ROWS=0 run strong1k_noedge --scaling strong --rows $((1024*N)) --steps 400 --warmup 40 --no-other-configs --e2e-steps 0
run weak --steps 20 --warmup 5 --no-other-configs --e2e-steps 0
timeout 300 python scripts/measure_configs.py c1 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   %-90s %8.2f Gcell/s  %.5f ms' % (d['config'][:90], d['gcell_per_s'], d['ms_per_step']))"
