#!/bin/bash
# gpu_call_ab.sh VAR A B: the -m gpu suite, then the default bench line (no baseline legs, no extra configs) with VAR=A and VAR=B in the same call
set -u
cd "$(dirname "$0")/.."
VAR=${1:-LLMREC_BRANCHES}; A=${2:-0}; B=${3:-1}
O=gpurun_out/r2ab
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for v in $A $B; do
  env $VAR=$v timeout 600 python bench.py --no-cpu --gpu-baseline 0 --extra 2 > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
j=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
print('$VAR=$v', j['ms_per_step'], j['e2e']['ms_per_step'], j['gpu_launches'], j['roofline']['families_ms'], j['roofline']['frac'], 'hoisted', [(v.get('ms_per_step'), v.get('error')) for v in j.get('configs', {}).values()])
PY
done
