#!/bin/bash
# Multi-GPU follow-up to tools/next_gpu_call.sh (box time is charged N x, so keep it short):
#   gpurun --gpus 2 --timeout 900  -- 'bash tools/next_gpu_call_multi.sh check'      # NCCL equality checks, world 2
#   gpurun --gpus 4 --timeout 1200 -- 'bash tools/next_gpu_call_multi.sh ab 4'       # item-sharded exchanges A/B, no 1-GPU base
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/next_gpu_call_multi.sh final 8'    # the reported 8-GPU line (with the same-call 1-GPU base)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/next_multi
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" bench.py --gpus "$1" "${@:3}"; }
case "${1:-check}" in
  check)
    timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu 2>&1 | tail -2
    LLMREC_TEST_EXPERIMENTAL=1 timeout 800 python -m pytest tests/test_experimental_gpu.py -q -m gpu -k "item_sharded or sharded_feature" 2>&1 | tail -3
    ;;
  ab)
    N=${2:-4}
    run $N 29601 --steps 10 --warmup 3 --n1-base 0 > $O/bench_${N}gpu_allreduce.json 2> $O/bench_${N}gpu_allreduce.err; cut -c1-300 $O/bench_${N}gpu_allreduce.json
    run $N 29602 --steps 10 --warmup 3 --n1-base 0 --item-sharded 1 > $O/bench_${N}gpu_item_sharded.json 2> $O/bench_${N}gpu_item_sharded.err; cut -c1-300 $O/bench_${N}gpu_item_sharded.json
    ;;
  final)
    N=${2:-8}
    run $N 29603 --steps 10 --warmup 3 ${3:-} > $O/bench_${N}gpu_final.json 2> $O/bench_${N}gpu_final.err; cut -c1-600 $O/bench_${N}gpu_final.json
    ;;
esac
