#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash scripts/profile.sh r02 --no-parity-check --no-other-configs > $OUT/r02_profile.log 2>&1
python scripts/ncu_summary.py $OUT/r02_march_raw.csv $OUT/r02_launches.csv > $OUT/r02_k_march_tma_lut.txt 2>&1; grep -E "time_duration|dram__bytes|registers|issue_active|long_scoreboard|traffic" $OUT/r02_k_march_tma_lut.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['ms_per_step'], d['parity_check']['ok'])"
