#!/bin/bash
# Eight ranks with the in-place peer halo read: strong and weak bench lines (each with its parity check), two sharded
# parity cases the bench does not cover (f64, WGSL semantics).
OUT=gpurun_out; mkdir -p $OUT
for mode in strong weak; do
  steps=400; [ $mode = weak ] && steps=100
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --scaling $mode --steps $steps --warmup 10 --no-other-configs > $OUT/r02_bench_${mode}_n8_pull.json 2> $OUT/r02_bench_${mode}_n8_pull.err
  echo "bench $mode rc=$?"
  python -c "
import json; d=json.loads([l for l in open('$OUT/r02_bench_${mode}_n8_pull.json') if l.startswith('{')][0]); print('$mode', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'] if d.get('e2e') else None, d['parity_check'])"
done
SBX_TEST_WORLDS=8 timeout 400 python -m pytest tests/test_gpu_multi.py -q -k "256-1024-40-1-peer or 256-1024-30-0-wgsl_peer or 64-128-90-0-peer_mix" > $OUT/r02_pytest_multi_pull_n8.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_multi_pull_n8.txt
