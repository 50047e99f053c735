#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_tile1; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tile.py -x -q -m gpu > $OUT/pytest_tile.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_tile.log
tail -30 $OUT/pytest_tile.log
echo "== tile" > $OUT/kbench.txt
timeout 120 python tools/kbench.py --dtypes f32 --reps 20 >> $OUT/kbench.txt 2>&1
echo "== one-shot" >> $OUT/kbench.txt
MSDA_TILE=0 timeout 120 python tools/kbench.py --dtypes f32 --reps 20 --kinds model >> $OUT/kbench.txt 2>&1
cat $OUT/kbench.txt
