#!/bin/bash
# round 2, call 1 (1 GPU): full GPU test suite, default bench line, PCIe ceilings, A/B of the arrive variant
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02_pytest_gpu.txt
tail -5 $OUT/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
tail -3 $OUT/r02_bench_n1.err; head -c 1500 $OUT/r02_bench_n1.json
timeout 300 python scripts/pcie_probe.py > $OUT/r02_pcie_probe_n1.txt 2>&1; cat $OUT/r02_pcie_probe_n1.txt
timeout 600 bash scripts/ab_variants.sh hl base sink0 > $OUT/r02_ab_arrive_sink.txt 2>&1; cat $OUT/r02_ab_arrive_sink.txt
