#!/bin/bash
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/r02b_pytest_multi_n$N.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02b_pytest_multi_n$N.txt
tail -30 $OUT/r02b_pytest_multi_n$N.txt
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "steps_per_frame or checkpoint" 2>&1 | tail -3
for halo in nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --halo $halo --no-other-configs > $OUT/r02b_bench_weak_${halo}_n$N.json 2> $OUT/r02b_bench_weak_${halo}_n$N.err
echo "bench $halo rc=$?"; tail -2 $OUT/r02b_bench_weak_${halo}_n$N.err; python -c "
import json; d=json.load(open('$OUT/r02b_bench_weak_${halo}_n$N.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity_check']['ok'])"
done
SBX_NO_BOUNDARY_FIRST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 --halo nccl --no-other-configs --e2e-steps 0 --no-parity-check > $OUT/r02b_bench_weak_nccl_serial_n$N.json 2> /dev/null
python -c "
import json; d=json.load(open('$OUT/r02b_bench_weak_nccl_serial_n$N.json')); print('nccl un-overlapped', d['value'], d['ms_per_step'])"
