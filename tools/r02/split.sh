#!/bin/bash
# r02: what bounds the backward? product vs REDs off / coarse REDs off / loads off / coarse off (MSDA_EXP 10..13)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_split; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
echo "== product" > $OUT/kbench.txt
python tools/kbench.py --dtypes f32 --reps 20 >> $OUT/kbench.txt 2>&1
for v in e10 e11 e12 e13; do
  echo "== $v" >> $OUT/kbench.txt
  MSDA_NATIVE_DIR=$V/$v python tools/kbench.py --dtypes f32 --kinds model --reps 20 >> $OUT/kbench.txt 2>&1
done
kill $SMI
cat $OUT/kbench.txt
for v in product e10 e12; do
  if [ $v = product ]; then D=tce-rvos_b200/_native; else D=$V/$v; fi
  MSDA_NATIVE_DIR=$D python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/plain_$v.log 2>&1 && \
  MSDA_NATIVE_DIR=$D ncu --set full --clock-control none -k regex:bwd_fast -s 2 -c 1 -o $OUT/prof_bwd_$v -f \
     python tools/kbench.py --dtypes f32 --kinds model --reps 2 > $OUT/ncu_$v.log 2>&1
  ncu -i $OUT/prof_bwd_$v.ncu-rep --page raw --csv > $OUT/prof_bwd_$v.raw.csv 2>/dev/null
done
ls -la $OUT
