#!/bin/bash
# round-2 profiling call (1 GPU): ncu --set full of the top kernels + launch lists; raw metric CSVs land in gpurun_out/r2p/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2p
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,smsp__cycles_active.avg"
echo "== projection kernels (netflix shape, grouped)"
REPS=1 timeout 600 ncu --metrics $M --clock-control none -k regex:"proj_(fwd|wgrad)_ts|wgrad_reduce|colsum" -s 8 -c 8 --csv --log-file $O/proj.csv python tools/prof_kernels.py proj > $O/proj.log 2>&1; tail -2 $O/proj.log
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"proj_wgrad_ts" -s 2 -c 1 -o $O/prof_wgrad_ts python tools/prof_kernels.py proj > /dev/null 2>&1
echo "== TMA-staged SpMM A/B"; LLMREC_SPMM_BULK=1 timeout 400 python tools/spmm_scale.py 1.0 2>&1 | tail -4
echo "== SpMM at the synthetic scale (both directions)"
timeout 900 ncu --metrics $M --clock-control none -k regex:"spmm_tile" -s 4 -c 4 --csv --log-file $O/spmm_syn.csv python tools/spmm_scale.py 1.0 > $O/spmm_syn.log 2>&1; tail -4 $O/spmm_syn.log
echo "== netflix-shape SpMM / fuse / bpr / score kernels inside one eager step + eval"
timeout 900 ncu --metrics $M --clock-control none -k regex:"spmm_|fuse_|bpr_|grad_init|adamw_kernel|score_topk|rescore" -s 150 -c 60 --csv --log-file $O/step_small.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --extra 0 --gpu-baseline 0 --min-seconds 0 --max-blocks 1 > $O/step_small.log 2>&1; tail -1 $O/step_small.log | cut -c1-200
echo "== launch list of one eager step (default engine)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv python tools/step_once.py --hoist 0 --steps 2 --spans 0 > $O/ncu_d.log 2>&1; tail -1 $O/ncu_d.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hoist.csv python tools/step_once.py --hoist 1 --steps 2 --spans 0 > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log
