#!/bin/bash
# r02: the remaining cells of the configs[4] sweep table (2 and 4 GPUs at 64 / 128 / 512 frames per GPU), final kernels
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_sweep4; mkdir -p $OUT
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
for n in 4 2; do for f in 64 128 512; do
  TR $n bench.py --gpus $n --workload sweep --frames $f > $OUT/sweep${f}_n$n.json 2> $OUT/sweep${f}_n$n.err; echo "n=$n f=$f rc=$?"; tail -n 1 $OUT/sweep${f}_n$n.json | cut -c1-200
done; done
