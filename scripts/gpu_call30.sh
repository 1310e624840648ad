#!/bin/bash
# Final single-GPU evidence of the round: GPU suite, launch list + full capture of the headline kernel, traffic.json tied to
# the kernel sources of this snapshot, then the bench line that reads it.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py > $OUT/r02_pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02_pytest_gpu_final.txt
bash scripts/profile.sh r02 --no-parity-check --no-other-configs > $OUT/r02_profile.log 2>&1
python scripts/ncu_summary.py $OUT/r02_march_raw.csv $OUT/r02_launches.csv > $OUT/r02_k_march_tma_lut.txt 2>&1
grep -E "time_duration|dram__bytes|registers|issue_active|long_scoreboard|traffic" $OUT/r02_k_march_tma_lut.txt
python scripts/update_traffic.py $OUT/r02_k_march_tma_lut.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['ms_per_step'], d['parity_check']['ok'])"
