#!/bin/bash
# gpu_call_multi.sh N [test]: the N-GPU bench line (and, with "test", the NCCL equality tests first)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r2m$N
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
if [ "${2:-}" = "test" ]; then timeout 1200 python -m pytest tests/test_dist_gpu.py -q -m gpu > $O/pytest_dist.log 2>&1; tail -3 $O/pytest_dist.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
tail -2 $O/bench.err | cut -c1-300
python - <<PY
import json
j=json.loads([l for l in open('$O/bench.json') if l.startswith('{')][-1])
print('N=$N', j['value'], j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], 'base', (j.get('same_workload_1gpu') or {}).get('ms_per_step'), 'speedup', j.get('speedup_vs_1gpu'), 'exch', j['config']['exchange_bytes_per_step'], 'eval', j['eval']['value'], 'roof', j['roofline']['frac'], j['roofline']['ms'])
PY
