#!/bin/bash
# round-2 call B (2 GPUs): NCCL equality tests of the sharded engines, then the N=2 bench line
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2b
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_dist_gpu.py -q -m gpu > $O/pytest_dist.log 2>&1; tail -3 $O/pytest_dist.log
LLMREC_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -q -m gpu -k "item_sharded or sharded_feature" > $O/pytest_exp.log 2>&1; tail -3 $O/pytest_exp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cut -c1-900 $O/bench_2gpu.json; tail -2 $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --item-sharded 1 --n1-base 0 > $O/bench_2gpu_is.json 2> $O/bench_2gpu_is.err; cut -c1-400 $O/bench_2gpu_is.json
