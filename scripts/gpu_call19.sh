#!/bin/bash
N=2
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "peer-2 or host-2 or foam-2 or 1536-256-10-0-nccl-2" > $OUT/r02d_pytest_multi_n2.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02d_pytest_multi_n2.txt
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $OUT/r02d_bench_${name}_n$N.json 2> $OUT/r02d_bench_${name}_n$N.err
  echo "bench $name rc=$?"; python -c "
import json; d=json.load(open('$OUT/r02d_bench_${name}_n$N.json')); print('$name', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'] and d['e2e']['ms_per_step'], d.get('parity_check',{}).get('ok'))
for o in d.get('other_configs') or []: print('   %-100s %9.4f ms %8.1f Gcell/s' % (o['config'][:100], o['ms_per_step'], o['gcell_per_s']))"; }
run weak --steps 20 --warmup 5
run weak_nccl --steps 20 --warmup 5 --halo nccl --no-other-configs --e2e-steps 0
