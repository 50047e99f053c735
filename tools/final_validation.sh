#!/bin/bash
# Round-end validation on a 1-GPU B200 box (run through gpurun): GPU tests, smoke, the contract benchmark on both
# arms (the driver's command line), the 1-GPU column of the configs[4] sweep, the ncu launch list of the benchmark
# command and one `--set full` capture per hot kernel.  The .ncu-rep files are converted to CSV pages on the box
# and removed (gpurun brings back at most 64 MiB).  Tag of the outputs: $1 (default r02).
set -x
T=${1:-r02}
O=gpurun_out/${T}_final; mkdir -p $O
python -m pytest tests -q -m gpu -rs > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -4 $O/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.json
python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-200 $O/bench_default.json
python bench.py --dtype bf16 --steps 100 --no-cpu-baseline > $O/bench_bf16.json 2>/dev/null
for f in 64 128 256 512; do python bench.py --workload sweep --frames $f --no-cpu-baseline > $O/bench_sweep$f.json 2> $O/bench_sweep$f.err; cut -c1-160 $O/bench_sweep$f.json; done
if [ -d tce-rvos_b200/_native_variants/e30 ]; then
  for v in e30 e31; do echo "== $v"; MSDA_NATIVE_DIR=tce-rvos_b200/_native_variants/$v python tools/kbench.py --dtypes f32 --reps 20; done > $O/kbench_evict_last.txt 2>&1
  echo "== product" >> $O/kbench_evict_last.txt; python tools/kbench.py --dtypes f32 --reps 20 >> $O/kbench_evict_last.txt 2>&1; cat $O/kbench_evict_last.txt
fi
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_launch.log 2>&1
cap() {  # name, kernel regex, launches, command...
  local name=$1 re=$2 n=$3; shift 3
  "$@" > $O/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$re -c $n -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/prof_$name.ncu-rep --page raw --csv > $O/prof_$name.raw.csv 2>/dev/null
  if [ "$name" = fwd_f32 ] || [ "$name" = bwd_f32 ]; then ncu -i $O/prof_$name.ncu-rep --page source --csv > $O/prof_$name.source.csv 2>/dev/null; fi
  rm -f $O/prof_$name.ncu-rep
}
cap fwd_f32 fwd_fast 2 python tools/kbench.py --kinds model --dtypes f32 --reps 1
cap bwd_f32 bwd_fast 2 python tools/kbench.py --kinds model --dtypes f32 --reps 1
du -sh $O; ls $O
