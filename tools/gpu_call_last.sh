#!/bin/bash
# shortest useful validation call: the golden-path and kernel GPU tests, then the default line + hoisted line without baseline legs
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2y
mkdir -p $O
timeout 150 python -m pytest tests/test_path_gpu.py tests/test_kernels_gpu.py -q -m gpu -x > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 100 python bench.py --no-cpu --gpu-baseline 0 --extra 2 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2y/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['e2e']['ms_per_step'], j['gpu_launches'], [(k, v.get('ms_per_step')) for k, v in j.get('configs', {}).items()])
PY
