#!/bin/bash
# final 1-GPU call of the round: the -m gpu suite, smoke(), the default bench line, the reference arm
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2z
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2z/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['families_ms'], j['roofline']['frac'], j['roofline']['traffic'])
print('parity', j.get('parity'))
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('error'))
PY
