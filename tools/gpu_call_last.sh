#!/bin/bash
# shortest A/B call: gpu_call_last.sh VAR A B -> the default line (no baseline legs, no extra configs) with VAR=A, VAR=B, VAR=A, VAR=B
set -u
cd "$(dirname "$0")/.."
VAR=${1:-LLMREC_PROF_LANE}; A=${2:-0}; B=${3:-1}
O=gpurun_out/r2y
mkdir -p $O
for v in $A $B $A $B; do
  env $VAR=$v timeout 60 python bench.py --no-cpu --gpu-baseline 0 --extra 0 --min-seconds 1 > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
j=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
print('$VAR=$v', j['ms_per_step'], j['e2e']['ms_per_step'], j['gpu_launches'])
PY
done
