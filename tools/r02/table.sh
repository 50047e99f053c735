#!/bin/bash
# r02: shared-memory point table (MSDA_TABLE=1, product) vs shuffle broadcast (MSDA_TABLE=0), and the full GPU suite
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_table; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
echo "== table (product)" > $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
echo "== shuffles (MSDA_TABLE=0)" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/notable timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
echo "== table, level set C 64 frames" >> $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 10 --levels C --frames 64 --kinds model --dtypes f32 >> $OUT/kbench.txt 2>&1
echo "== shuffles, level set C 64 frames" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/notable timeout 200 python tools/kbench.py --reps 10 --levels C --frames 64 --kinds model --dtypes f32 >> $OUT/kbench.txt 2>&1
cat $OUT/kbench.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
