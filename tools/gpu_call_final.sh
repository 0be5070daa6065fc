#!/bin/bash
# final 1-GPU call of the round: the -m gpu suite, smoke(), the default bench line, the reference arm
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2z
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-300 $O/bench_ref.json
echo "== same call, branches off (A/B of the graph-branch schedule)"
LLMREC_BRANCHES=0 timeout 200 python bench.py --no-cpu --gpu-baseline 0 --extra 2 > $O/bench_nobranch.json 2> $O/bench_nobranch.err
echo "== launch lists (ncu serialises launches: shares only)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv python tools/step_once.py --hoist 0 --steps 2 --spans 0 > $O/ncu_d.log 2>&1; tail -1 $O/ncu_d.log | cut -c1-160
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hoist.csv python tools/step_once.py --hoist 1 --steps 2 --spans 0 > $O/ncu_h.log 2>&1; tail -1 $O/ncu_h.log | cut -c1-160
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,smsp__cycles_active.avg"
echo "== scoring kernels (netflix eval leg)"
timeout 240 ncu --metrics $M --clock-control none -k regex:"score_topk|rescore" -c 6 --csv --log-file $O/score.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --extra 0 --gpu-baseline 0 --min-seconds 0 --max-blocks 1 > $O/score.log 2>&1; tail -1 $O/score.log | cut -c1-120
python - <<'PY'
import json
for f in ('bench_nobranch.json',):
    try:
        j=json.loads([l for l in open('gpurun_out/r2z/'+f) if l.startswith('{')][-1])
        print(f, j['ms_per_step'], j['e2e']['ms_per_step'], [(k, v.get('ms_per_step')) for k, v in j.get('configs', {}).items()])
    except Exception as e: print(f, 'unreadable', e)
j=json.loads([l for l in open('gpurun_out/r2z/bench.json') if l.startswith('{')][-1])
print('default', j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['families_ms'], j['roofline']['frac'], j['roofline']['traffic'])
print('parity', j.get('parity'))
for k,v in j['configs'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('error'))
PY
