#!/bin/bash
# two GPUs: latitude-band sharding tests (peer push + NCCL) and the N=2 bench line
set -u
O=gpurun_out; mkdir -p $O
echo "== multi-GPU tests"; timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 | tee $O/pytest_multi.txt
echo "== bench N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench_n2.json 2> $O/bench_n2.err; cat $O/bench_n2.json; tail -3 $O/bench_n2.err
