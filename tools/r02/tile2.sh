#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_tile2; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
KB="timeout 120 python tools/kbench.py --dtypes f32 --kinds model --reps 20"
echo "== product (spin lock)" > $OUT/kbench.txt; $KB >> $OUT/kbench.txt 2>&1
for v in t_lock0 t_lock2 t_e20; do echo "== $v" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/$v $KB >> $OUT/kbench.txt 2>&1; done
echo "== product cap=0 (persistent schedule only)" >> $OUT/kbench.txt; MSDA_TILE_CAP=0 $KB >> $OUT/kbench.txt 2>&1
echo "== product cap=60 (level 3 only)" >> $OUT/kbench.txt; MSDA_TILE_CAP=60 $KB >> $OUT/kbench.txt 2>&1
for b in 1 2 8 16; do echo "== product batch=$b" >> $OUT/kbench.txt; MSDA_TILE_BATCH=$b $KB >> $OUT/kbench.txt 2>&1; done
cat $OUT/kbench.txt
timeout 120 python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 4 -c 2 -o $OUT/prof_tile -f \
   python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof_tile.ncu-rep --page raw --csv > $OUT/prof_tile.raw.csv 2>/dev/null
ncu -i $OUT/prof_tile.ncu-rep --page source --csv > $OUT/prof_tile.source.csv 2>/dev/null
tail -3 $OUT/ncu.log
