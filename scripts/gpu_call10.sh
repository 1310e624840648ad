#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4
timeout 900 bash scripts/ab_variants.sh hl base nosplit head > $OUT/r02_ab_split.txt 2>&1; cat $OUT/r02_ab_split.txt
timeout 300 python scripts/resident_frame_probe.py > $OUT/r02_resident_frame_probe.txt 2>&1; cat $OUT/r02_resident_frame_probe.txt
