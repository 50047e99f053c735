#!/bin/bash
# r02: query-major item map (a CTA = 8 consecutive queries of one head group) vs item-major (4 queries x 2 groups)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_qmap; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
echo "== product (item-major)" > $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 --dtypes f32 >> $OUT/kbench.txt 2>&1
for v in qf qb; do echo "== $v" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/$v timeout 200 python tools/kbench.py --reps 20 --dtypes f32 >> $OUT/kbench.txt 2>&1; done
cat $OUT/kbench.txt
for v in qf qb; do MSDA_NATIVE_DIR=$V/$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config2_fp32 or stress or decoder or overlapping" 2>&1 | tail -2; done
