#!/bin/bash
# round-2 call A (1 GPU): the -m gpu suite, the full default bench line, and the launch list of an eager step
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2a
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-600 $O/bench.json; tail -3 $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --graph 0 --extra 0 --gpu-baseline 0 --min-seconds 0 > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log
