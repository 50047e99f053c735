#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_bench2; mkdir -p $OUT
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_config2.json 2> $OUT/bench_config2.err; echo "rc=$?"
python bench.py --workload model3 --steps 3 > $OUT/bench_model3.json 2> $OUT/bench_model3.err; echo "rc=$?"; tail -c 1500 $OUT/bench_model3.err
python bench.py --workload model4 --steps 3 > $OUT/bench_model4.json 2> $OUT/bench_model4.err; echo "rc=$?"; tail -c 1500 $OUT/bench_model4.err
for f in $OUT/*.json; do echo "== $f"; tail -n 1 $f | cut -c1-1500; done
