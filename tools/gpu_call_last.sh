#!/bin/bash
# shortest sanity call: the golden-path GPU tests on the current tree
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r2y
timeout 80 python -m pytest tests/test_path_gpu.py -q -m gpu -x > gpurun_out/r2y/pytest_path.log 2>&1; tail -2 gpurun_out/r2y/pytest_path.log
