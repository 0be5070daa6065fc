#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2i
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
echo "== two accumulators per tile (DBG=8) vs default: timing + correctness"
timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
LLMREC_PROJ_DBG=8 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
LLMREC_PROJ_DBG=8 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "proj" 2>&1 | tail -2
echo "== synthetic 1-GPU step with the TMA-staged SpMM on the user-table gathers"
timeout 600 python bench.py --workload synthetic --steps 5 --warmup 3 --min-seconds 1 --eval-users 20480 > $O/syn1.json 2> $O/syn1.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2i/syn1.json') if l.startswith('{')][-1])
print('syn 1gpu', j['ms_per_step'], j['config']['timing'], 'roof', j['roofline']['ms'], j['roofline']['frac'], j['roofline']['gather_bound_gbs'])
PY
