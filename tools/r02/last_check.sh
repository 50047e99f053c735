#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/r02_last; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-160 $OUT/bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_last/bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02_last/bench_reference.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'parity',d['shard_parity']['ok'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],'ref',r['value'],r['cpu_baseline']['cores'],'ratio',d['value']/r['value'],'e2e ratio',d['e2e']['value']/r['value'])
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
