#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2d
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python tools/step_once.py --hoist 1 --steps 5 2>/dev/null | tail -14
timeout 300 python tools/step_once.py --hoist 0 --steps 5 2>/dev/null | tail -14
timeout 600 python bench.py --no-cpu --gpu-baseline 0 > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2d/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['roofline']['families_ms'])
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), v.get('error'))
PY
echo "== W-refetch experiment (timing only): projection forward with and without the W tile loads"
timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
LLMREC_PROJ_SKIPW=1 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
