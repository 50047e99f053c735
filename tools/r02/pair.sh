#!/bin/bash
# r02: two-query warps with cross-query merging of the reductions (MSDA_PAIR=1) vs product
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_pair; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
MSDA_NATIVE_DIR=$V/pair timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or config2 or overlapping or stress or non_finite or reference_cuda or decoder or autograd or sweep_size or random" > $OUT/pytest_pair.log 2>&1; echo "rc=$?" >> $OUT/pytest_pair.log; tail -6 $OUT/pytest_pair.log
echo "== pair" > $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/pair timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
echo "== product" >> $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 --dtypes f32 --kinds model >> $OUT/kbench.txt 2>&1
cat $OUT/kbench.txt
