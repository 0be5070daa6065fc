#!/bin/bash
# quick 1-GPU check: the -m gpu suite and the default bench line without the CPU / GPU-torch baseline legs
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2q
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py --no-cpu --gpu-baseline 0 > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err | cut -c1-200
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2q/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['e2e']['ms_per_step'], j['gpu_launches'], j['roofline']['families_ms'], j['roofline']['frac'])
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('families_ms'))
PY
