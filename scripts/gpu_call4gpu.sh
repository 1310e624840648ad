#!/bin/bash
N=4
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q \
  -k "128-120-0-peer-4 or 1024-60-0-peer-4 or 1300-256-12-0-host-4 or 128-40-0-foam-4 or 1024-40-1-peer-4 or 1024-30-0-wgsl_peer-4 or 1536-256-10-0-nccl-4 or 128-12-0-wgsl_nccl-4 or 128-140-2-nccl-4" \
  > $OUT/r02_pytest_multi_n4.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02_pytest_multi_n4.txt; tail -4 $OUT/r02_pytest_multi_n4.txt
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $OUT/r02_bench_${name}_n$N.json 2> $OUT/r02_bench_${name}_n$N.err
  echo "bench $name rc=$?"; tail -2 $OUT/r02_bench_${name}_n$N.err | cut -c1-300; python -c "
import json; d=json.load(open('$OUT/r02_bench_${name}_n$N.json')); print('$name', d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['ms_per_step'], d.get('parity_check',{}).get('ok'))
for o in d.get('other_configs') or []: print('   %-100s %9.4f ms %8.1f Gcell/s' % (o['config'][:100], o['ms_per_step'], o['gcell_per_s']))"; }
run weak --steps 20 --warmup 5
run foam --workload foam --steps 20 --warmup 5 --e2e-steps 2
