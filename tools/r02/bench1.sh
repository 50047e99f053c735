#!/bin/bash
# r02: new bench.py (CUDA-graph timed region, shard parity, fixed CPU sample), sweep workload, full GPU suite
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_bench1; mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/bench_config2.json 2> $OUT/bench_config2.err; echo "rc=$?"; tail -c 600 $OUT/bench_config2.err
python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > $OUT/bench_config2_eager.json 2>> $OUT/bench_config2.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python bench.py --workload sweep --frames 64 > $OUT/bench_sweep64.json 2> $OUT/bench_sweep64.err; tail -c 600 $OUT/bench_sweep64.err
python bench.py --workload sweep --frames 1024 --no-cpu-baseline > $OUT/bench_sweep1024.json 2> $OUT/bench_sweep1024.err; tail -c 600 $OUT/bench_sweep1024.err
python bench.py --dtype bf16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_config2_bf16.json 2> $OUT/bench_config2_bf16.err
for f in $OUT/*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("unparsable", e); sys.exit(0)
r=d.get("roofline") or {}
print({k:d.get(k) for k in ("value","ms_per_step","ms_per_step_eager","gpu_launches")}, "frac",r.get("frac"), (r.get("passes") or {}).get("forward",{}).get("us"), (r.get("passes") or {}).get("backward",{}).get("us"))
print(" e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"), "parity", d.get("shard_parity",{}).get("ok"), d.get("shard_parity",{}).get("max_err_over_ranks"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), (d.get("cpu_baseline") or {}).get("cores"), "clocks", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
PY
done
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
