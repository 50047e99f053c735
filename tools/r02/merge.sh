#!/bin/bash
# r02: duplicate-corner merging inside a (query, head, level): MSDA_MERGE=1 (product) vs 0
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_merge; mkdir -p $OUT
V=tce-rvos_b200/_native_variants
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overlapping or golden or config2 or stress or non_finite or reference_cuda or prologue" > $OUT/pytest_sel.log 2>&1; echo "rc=$?" >> $OUT/pytest_sel.log; tail -5 $OUT/pytest_sel.log
echo "== merge (product)" > $OUT/kbench.txt; timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
echo "== no merge (MSDA_MERGE=0; 5-shuffle hand-over without folding)" >> $OUT/kbench.txt; MSDA_NATIVE_DIR=$V/nomerge timeout 200 python tools/kbench.py --reps 20 >> $OUT/kbench.txt 2>&1
cat $OUT/kbench.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -n 1 $OUT/bench.json | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
