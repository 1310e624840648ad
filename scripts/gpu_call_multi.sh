#!/bin/bash
# multi-GPU call: sharded parity tests (world <= visible GPUs), bench lines (weak / strong / foam), PCIe ceilings
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/r02_pytest_multi_n$N.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02_pytest_multi_n$N.txt
tail -4 $OUT/r02_pytest_multi_n$N.txt
run() { # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $OUT/r02_bench_${name}_n$N.json 2> $OUT/r02_bench_${name}_n$N.err
  echo "bench $name rc=$?"; tail -2 $OUT/r02_bench_${name}_n$N.err; head -c 600 $OUT/r02_bench_${name}_n$N.json; echo
}
run weak --steps 20 --warmup 5
run strong --scaling strong --steps 200 --warmup 20 --no-other-configs
run foam --workload foam --steps 20 --warmup 5
run weak_nccl --steps 20 --warmup 5 --halo nccl --no-other-configs
timeout 300 python scripts/pcie_probe.py > $OUT/r02_pcie_probe_n$N.txt 2>&1; cat $OUT/r02_pcie_probe_n$N.txt
