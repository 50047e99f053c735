#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_f3; mkdir -p $OUT
python tools/smallcall_bench.py > $OUT/smallcall.txt 2>&1; cat $OUT/smallcall.txt
python tools/model_dropin.py --config 3 --steps 3 --fused-topk --compare-topk > $OUT/model3_fused_topk.json 2> $OUT/model3_fused_topk.err; echo "rc=$?"; tail -c 800 $OUT/model3_fused_topk.err; tail -n 1 $OUT/model3_fused_topk.json | cut -c1-2500
python tools/model_dropin.py --config 3 --steps 3 > $OUT/model3_plain.json 2> $OUT/model3_plain.err; tail -n 1 $OUT/model3_plain.json | cut -c1-600
python tools/model_dropin.py --config 3 --steps 3 --fused-topk --cuda-graph > $OUT/model3_graph.json 2> $OUT/model3_graph.err; echo "rc=$?"; tail -c 800 $OUT/model3_graph.err; tail -n 1 $OUT/model3_graph.json | cut -c1-2500
