#!/bin/bash
# r02: persistent backward, proportional-backoff ticket lock + predicate-free section
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_tile4; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
KB="timeout 120 python tools/kbench.py --dtypes f32 --kinds model --reps 20"
timeout 300 python -m pytest tests/test_gpu_tile.py -x -q -m gpu > $OUT/pytest_tile.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_tile.log
tail -5 $OUT/pytest_tile.log
echo "== product (16 warps, sleep 100)" > $OUT/kbench.txt; $KB >> $OUT/kbench.txt 2>&1
for v in t_w24 t_w24s50 t_w16s200; do echo "== $v" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/$v $KB >> $OUT/kbench.txt 2>&1; done
echo "== product cap=60 (level 3 only)" >> $OUT/kbench.txt; MSDA_TILE_CAP=60 $KB >> $OUT/kbench.txt 2>&1
cat $OUT/kbench.txt
timeout 120 python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bwd_tile_kernel -s 2 -c 1 -o $OUT/prof_tile -f \
   python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof_tile.ncu-rep --page raw --csv > $OUT/prof_tile.raw.csv 2>/dev/null
ncu -i $OUT/prof_tile.ncu-rep --page source --csv > $OUT/prof_tile.source.csv 2>/dev/null
rm -f $OUT/prof_tile.ncu-rep
tail -3 $OUT/ncu.log
