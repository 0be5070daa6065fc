#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2h
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err | cut -c1-300; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2h/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['families_ms'], j['roofline']['frac'])
print('parity', j.get('parity')); print('cpu', j.get('cpu_baseline',{}).get('value'), 'gpu_torch', j.get('gpu_torch_baseline',{}).get('ms_per_step'))
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('error'))
PY
