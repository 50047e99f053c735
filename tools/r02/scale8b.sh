#!/bin/bash
# r02: config2 at 1/2/4/8 GPUs with the final kernels (backward merge), one 8xB200 box
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_scale8b; mkdir -p $OUT
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
run() { name=$1; shift; echo "== $name"; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "rc=$?"; tail -n 1 $OUT/$name.json | cut -c1-300; }
run config2_n1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
for n in 2 4 8; do run config2_n$n TR $n bench.py --gpus $n --steps 20 --warmup 5; done
run sweep1024_n8 TR 8 bench.py --gpus 8 --workload sweep --frames 1024
