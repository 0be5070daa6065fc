#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2f
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
echo "== v2 forward skeleton experiments (timing only; wrong results): bit 1 = no W loads, 2 = no MMAs, 4 = no transform"
for v in 0 1 2 4 6 7; do echo "-- SKIPW=$v"; LLMREC_PROJ_SKIPW=$v LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=2 timeout 200 python tools/prof_kernels.py proj 2>&1 | grep proj_fwd; done
echo "-- MODE=1 (v1 plain TF32)"; MODE=1 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
echo "== hoisted step launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hoist.csv python tools/step_once.py --hoist 1 --steps 2 --spans 0 > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log
echo "== synthetic step launch list (scale 0.5)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_syn.csv python bench.py --workload synthetic --syn-scale 0.5 --steps 2 --warmup 3 --min-seconds 0 --max-blocks 1 --eval-users 2048 > $O/ncu_syn.log 2>&1; tail -1 $O/ncu_syn.log | cut -c1-300
