#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2g
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
echo "== many-tiles-per-CTA projection test: v2 / v3 fwd / v3 wgrad"
for v in "LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=2" "LLMREC_PROJ_FWD_V=3 LLMREC_PROJ_WG_V=2" "LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=3"; do
  echo "-- $v"; env $v CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k many_tiles 2>&1 | tail -4 | cut -c1-300
done
echo "== projection kernels alone: v3 ring splits"
for nw in 3 4 6 8; do echo "-- NW=$nw"; LLMREC_PROJ_NW=$nw timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2; done
LLMREC_PROJ_FWD_V=2 LLMREC_PROJ_WG_V=2 timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
