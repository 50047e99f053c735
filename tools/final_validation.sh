#!/bin/bash
# Round-end validation on a B200 box (run through gpurun): GPU tests, smoke, the contract benchmark on both
# arms, the ncu launch list of the benchmark command and one `--set full` capture per kernel.  The .ncu-rep
# files are converted to CSV pages on the box and removed (gpurun brings back at most 64 MiB).
set -x
O=gpurun_out
python -m pytest tests -q -m gpu -rs > $O/pytest_gpu_final.log 2>&1; tail -4 $O/pytest_gpu_final.log
python __graft_entry__.py smoke > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
python bench.py > $O/bench_final.log 2> $O/bench_final.err; tail -c 300 $O/bench_final.log
python bench.py --dtype bf16 --steps 100 --no-cpu-baseline > $O/bench_final_bf16.log 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_ref.log 2>/dev/null; cut -c1-160 $O/bench_final_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
cap() {  # name, kernel regex, launches, command...
  local name=$1 re=$2 n=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$re -c $n -f -o $O/prof_r1_$name "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/prof_r1_$name.ncu-rep --page raw --csv > $O/prof_r1_$name.raw.csv 2>/dev/null
  if [ "$name" = fwd_f32 ] || [ "$name" = bwd_f32 ]; then ncu -i $O/prof_r1_$name.ncu-rep --page source --csv > $O/prof_r1_$name.source.csv 2>/dev/null; fi
  rm -f $O/prof_r1_$name.ncu-rep
}
for dt in f32 bf16; do
  cap fwd_$dt fwd_fast 2 python tools/kbench.py --kinds model --dtypes $dt --reps 1
  cap bwd_$dt bwd_fast 2 python tools/kbench.py --kinds model --dtypes $dt --reps 1
done
AUXBENCH_NO_GRAPH=1 cap value_epilogue value_epilogue 8 python tools/auxbench.py
AUXBENCH_NO_GRAPH=1 cap topk topk_samples 4 python tools/auxbench.py
du -sh $O; ls -la $O | head -40
