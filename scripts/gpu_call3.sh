#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02b_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02b_pytest_gpu.txt
tail -15 $OUT/r02b_pytest_gpu.txt
timeout 600 bash scripts/ab_variants.sh exact3 base exact3f64 > $OUT/r02b_ab_exact3.txt 2>&1; cat $OUT/r02b_ab_exact3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02b_bench_n1.json 2> $OUT/r02b_bench_n1.err; echo "bench rc=$?"; tail -3 $OUT/r02b_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench_n1.json'))
print(d['value'], d['e2e']['ms_per_step'], d['parity_check']['ok'])
for o in d['other_configs']: print("%-100s %9.4f ms %8.1f Gcell/s %5.1f%%"%(o['config'][:100],o['ms_per_step'],o['gcell_per_s'],100*o['frac_of_measured_hbm']))
PY
