#!/bin/bash
# r02: 1/2/4/8-GPU runs of bench.py on one 8xB200 box: config2, sweep, whole-model arms, concurrent-copy probe
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_scale8; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
run() { name=$1; shift; echo "== $name"; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "rc=$?"; tail -n 1 $OUT/$name.json | cut -c1-400; }
run config2_n1 python bench.py --gpus 1 --steps 20 --warmup 5
for n in 2 4 8; do run config2_n$n TR $n bench.py --gpus $n --steps 20 --warmup 5; done
run ref_n2 TR 2 bench.py --impl reference --gpus 2 --steps 20 --warmup 5
for f in 64 128 256 512 1024; do run sweep${f}_n8 TR 8 bench.py --gpus 8 --workload sweep --frames $f; done
for n in 2 4; do for f in 256 1024; do run sweep${f}_n$n TR $n bench.py --gpus $n --workload sweep --frames $f; done; done
for n in 2 4 8; do run model4_n$n TR $n bench.py --gpus $n --workload model4 --steps 5; done
run model3_n8 TR 8 bench.py --gpus 8 --workload model3 --steps 5
run copy_n1 python tools/copy_probe.py
for n in 2 4 8; do run copy_n$n TR $n tools/copy_probe.py; done
