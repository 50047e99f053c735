#!/bin/bash
# r02: does it matter WHEN a warp issues its reductions?  EXP 14 = all 64 before the corner loads, EXP 15 = all after the
# gradients are stored (both: +16 shuffles per item), product = interleaved per point
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_redorder; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
echo "== product (interleaved)" > $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 --dtypes f32 >> $OUT/kbench.txt 2>&1
for v in e14 e15; do echo "== $v" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/$v timeout 200 python tools/kbench.py --reps 20 --dtypes f32 >> $OUT/kbench.txt 2>&1; done
cat $OUT/kbench.txt
for v in e14 e15; do MSDA_NATIVE_DIR=$V/$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_fp32 or config2_fp32 or overlapping" 2>&1 | tail -1; done
