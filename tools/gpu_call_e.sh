#!/bin/bash
# round-2 call E (1 GPU): v3 projection kernels (decoupled rings) vs v2, TMA-staged SpMM A/B at the synthetic scale, hoisted launch list
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2e
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
echo "== projection tests on the v3 kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_path_gpu.py -q -m gpu -x > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo "== projection kernels alone: v3 (default) vs v2"
timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=2 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
D=128 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
D=128 LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=2 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
echo "== SpMM at the synthetic scale: register gathers vs TMA-staged rows"
LLMREC_SPMM_BULK=0 timeout 400 python tools/spmm_scale.py 1.0 2>&1 | tail -4
LLMREC_SPMM_BULK=1 timeout 400 python tools/spmm_scale.py 1.0 2>&1 | tail -4
echo "== hoisted step launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hoist.csv python tools/step_once.py --hoist 1 --steps 2 --spans 0 > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log
timeout 600 python bench.py --no-cpu --gpu-baseline 0 > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2e/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['roofline']['families_ms'], j['roofline']['frac'])
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('families_ms'))
PY
