#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== bench"; timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
echo "== configs"; timeout 600 python scripts/measure_configs.py > $O/configs.txt 2> $O/configs.err; tail -28 $O/configs.txt
echo "== ncu"; timeout 400 scripts/profile.sh r01d > $O/profile.log 2>&1; tail -3 $O/profile.log
