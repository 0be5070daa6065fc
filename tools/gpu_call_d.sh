#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2d
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_path_gpu.py -q -m gpu -k "mask_dropout or full_test_flag" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python tools/step_once.py --hoist 1 --steps 5 2>/dev/null | tail -14
timeout 300 python tools/step_once.py --hoist 0 --steps 5 2>/dev/null | tail -14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hoist.csv python tools/step_once.py --hoist 1 --steps 2 --spans 0 > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv python tools/step_once.py --hoist 0 --steps 2 --spans 0 > $O/ncu_d.log 2>&1; tail -1 $O/ncu_d.log
