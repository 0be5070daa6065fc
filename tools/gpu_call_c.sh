#!/bin/bash
# round-2 call C (1 GPU): -m gpu suite (hoisted engine, demand-driven step, AUC), then the default bench line
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
