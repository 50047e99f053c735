#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_merge2; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
echo "== product (bwd merge, fwd separable hand-over; bwd 3 CTAs/SM)" > $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
for v in mb2 mb4; do echo "== $v" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/$v timeout 200 python tools/kbench.py --reps 20 --dtypes f32 >> $OUT/kbench.txt 2>&1; done
cat $OUT/kbench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overlapping or golden or config2 or stress or non_finite or reference_cuda or prologue" > $OUT/pytest_sel.log 2>&1; echo "rc=$?" >> $OUT/pytest_sel.log; tail -3 $OUT/pytest_sel.log
python tools/kbench.py --kinds model --dtypes f32 --reps 1 > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bwd_fast -c 1 -f -o $OUT/prof_bwd python tools/kbench.py --kinds model --dtypes f32 --reps 1 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof_bwd.ncu-rep --page raw --csv > $OUT/prof_bwd.raw.csv 2>/dev/null; rm -f $OUT/prof_bwd.ncu-rep
