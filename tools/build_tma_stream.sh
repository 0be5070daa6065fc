#!/bin/bash
# Builds tools/tma_stream against the in-tree library (tensor-map encoder, error string).
set -e
cd "$(dirname "$0")/.."
python -m llmrec_b200.build > /dev/null
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I llmrec_b200/csrc \
  -o tools/tma_stream tools/tma_stream.cu -L llmrec_b200/lib -lllmrec_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../llmrec_b200/lib'
echo tools/tma_stream
