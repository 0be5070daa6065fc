#!/bin/bash
# First GPU call after a round that ended without GPU budget: validates everything that was written on the CPU side and
# produces the A/B numbers DESIGN.md 7 asks for, all on ONE box (box-to-box variance is +-30 %, only same-call numbers compare).
#   gpurun --timeout 1800 -- 'bash tools/next_gpu_call.sh'     (about 18 minutes of box time)
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
echo "== k-rotation (LLMREC_PROJ_KROT=1): concurrent CTAs read different feature columns"
LLMREC_PROJ_KROT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" > $O/pytest_krot.log 2>&1; tail -1 $O/pytest_krot.log
LLMREC_PROJ_KROT=1 timeout 400 python bench.py --no-cpu > $O/bench_rows_krot.json 2> $O/bench_rows_krot.err
LLMREC_PROJ_KROT=1 timeout 400 python bench.py --no-cpu --feat_layout panels > $O/bench_panels_krot.json 2> $O/bench_panels_krot.err
echo "== one TMA box for W_hi|W_lo (LLMREC_PROJ_WBOX=1)"
LLMREC_PROJ_WBOX=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" > $O/pytest_wbox.log 2>&1; tail -1 $O/pytest_wbox.log
LLMREC_PROJ_WBOX=1 timeout 400 python bench.py --no-cpu > $O/bench_rows_wbox.json 2> $O/bench_rows_wbox.err
echo "== rank-3 TMA boxes in proj_wgrad: X tile (needs panels) / dY tile"
LLMREC_PROJ_X3D=1 LLMREC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu -k "panel_layout_equals" > $O/pytest_x3d.log 2>&1; tail -1 $O/pytest_x3d.log
LLMREC_PROJ_G3D=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "projection" > $O/pytest_g3d.log 2>&1; tail -1 $O/pytest_g3d.log
LLMREC_PROJ_X3D=1 timeout 400 python bench.py --no-cpu --feat_layout panels > $O/bench_panels_x3d.json 2> $O/bench_panels_x3d.err
LLMREC_PROJ_G3D=1 timeout 400 python bench.py --no-cpu > $O/bench_rows_g3d.json 2> $O/bench_rows_g3d.err
LLMREC_PROJ_X3D=1 LLMREC_PROJ_G3D=1 LLMREC_PROJ_WBOX=1 timeout 400 python bench.py --no-cpu --feat_layout panels > $O/bench_panels_all.json 2> $O/bench_panels_all.err
python - <<'PY'
import json
for name in ("rows", "panels", "rows_krot", "panels_krot", "rows_wbox", "panels_x3d", "rows_g3d", "panels_all"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/next/bench_{name}.json") if l.startswith("{")][0])
        print(name, "ms/step", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "families", j["roofline"]["families_ms"], "eval users/s", j.get("eval", {}).get("value"))
    except Exception as e:
        print(name, "no line:", e)
PY
