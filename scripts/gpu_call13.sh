#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
bash scripts/profile.sh r02 --no-parity-check --no-other-configs > $OUT/r02_profile.log 2>&1
python scripts/ncu_summary.py $OUT/r02_march_raw.csv $OUT/r02_launches.csv > $OUT/r02_k_march_tma_lut.txt 2>&1; tail -45 $OUT/r02_k_march_tma_lut.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['parity_check']['ok'], d['cpu_baseline']['value'])
for o in d['other_configs']: print("%-100s %9.4f ms %8.1f Gcell/s %5.1f%%"%(o['config'][:100],o['ms_per_step'],o['gcell_per_s'],100*o['frac_of_measured_hbm']))
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r02_bench_reference_arm.json 2>/dev/null; head -c 400 $OUT/r02_bench_reference_arm.json
