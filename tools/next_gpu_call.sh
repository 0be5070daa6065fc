#!/bin/bash
# First GPU call after a round that ended without GPU budget: validates everything that was written on the CPU side and
# produces the A/B numbers DESIGN.md 7 asks for, all on ONE box (box-to-box variance is +-30 %, only same-call numbers compare).
#   gpurun --timeout 1800 -- 'bash tools/next_gpu_call.sh'     (about 12 minutes of box time)
# Outputs land in gpurun_out/next/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/next
mkdir -p $O
python -m llmrec_b200.build > $O/build.log 2>&1
echo "== validated suite" ;      timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
echo "== experimental suite" ;   LLMREC_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -q -m gpu > $O/pytest_experimental.log 2>&1; tail -3 $O/pytest_experimental.log
echo "== fetch-pattern ceilings"; bash tools/build_tma_stream.sh > /dev/null 2>&1 && timeout 400 ./tools/tma_stream > $O/tma_stream.txt 2>&1; cat $O/tma_stream.txt
echo "== bench rows";            timeout 400 python bench.py --no-cpu > $O/bench_rows.json 2> $O/bench_rows.err; cut -c1-400 $O/bench_rows.json
echo "== bench panels";          timeout 400 python bench.py --no-cpu --feat_layout panels > $O/bench_panels.json 2> $O/bench_panels.err; cut -c1-400 $O/bench_panels.json
echo "== projection kernels alone, every fetch variant (tools/prof_kernels.py proj: ms and GB/s of the grouped fwd / wgrad launches)"
for v in "" "PANELS=1" "LLMREC_PROJ_KROT=1" "PANELS=1 LLMREC_PROJ_KROT=1" "LLMREC_PROJ_WBOX=1" "PANELS=1 LLMREC_PROJ_X3D=1" "LLMREC_PROJ_G3D=1" \
         "PANELS=1 LLMREC_PROJ_X3D=1 LLMREC_PROJ_WBOX=1" "PANELS=1 LLMREC_PROJ_X3D=1 LLMREC_PROJ_WBOX=1 LLMREC_PROJ_KROT=1" "MODE=1" "MODE=1 PANELS=1" "LLMREC_WG_ROWS=1024" "LLMREC_WG_ROWS=4096"; do
  echo "-- ${v:-default}"; env $v timeout 200 python tools/prof_kernels.py proj 2>&1 | tail -2
done | tee $O/prof_variants.txt
echo "== netflix-shaped SpMM launches per variant (more gathers in flight / more warps per SM)"
for v in 0 1 2 3 4; do echo "-- LLMREC_SPMM_VARIANT=$v"; LLMREC_SPMM_VARIANT=$v timeout 200 python tools/prof_kernels.py spmm 2>&1 | grep "tile=0"; done | tee $O/spmm_variants.txt
echo "== correctness of the variants"
LLMREC_PROJ_KROT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" 2>&1 | tail -1
LLMREC_PROJ_WBOX=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" 2>&1 | tail -1
LLMREC_PROJ_G3D=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" 2>&1 | tail -1
LLMREC_PROJ_X3D=1 LLMREC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu -k "panel_layout_equals" 2>&1 | tail -1
python - <<'PY'
import json
for name in ("rows", "panels"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/next/bench_{name}.json") if l.startswith("{")][0])
        print(name, "ms/step", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "families", j["roofline"]["families_ms"], "eval users/s", j.get("eval", {}).get("value"))
    except Exception as e:
        print(name, "no line:", e)
PY
echo "== SpMM at the 10M x 1M x 200M synthetic scale: gather variants and column windows (1 GPU)"
for v in 0 1 3; do LLMREC_SPMM_VARIANT=$v COLWIN=$([ $v = 0 ] && echo 2,4,8) timeout 400 python tools/spmm_scale.py 1.0 2>&1 | tail -12; done | tee $O/spmm_scale.txt
