"""A few EAGER training steps of one bench configuration (kernel names visible to ncu / KernelTimer spans):
    python tools/step_once.py [--workload netflix] [--hoist 1] [--steps 3] [--spans 1]
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv python tools/step_once.py ..."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="netflix")
ap.add_argument("--hoist", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--spans", type=int, default=1)
a = ap.parse_args()
ba = argparse.Namespace(proj_mode="3xtf32", host_sampler="native", graph=0)
tr, gen, args = bench.make_trainer(a.workload, ba, extra=(["--hoist_side", "1"] if a.hoist else []))
hp = tr.hot
for _ in range(3):
    tr.train_next_batch()
torch.cuda.synchronize()
if a.spans:
    from llmrec_b200.engine import KernelTimer
    hp.timer = KernelTimer()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    tr.train_next_batch()
e1.record()
torch.cuda.synchronize()
print(f"eager: {e0.elapsed_time(e1) / a.steps:.4f} ms/step over {a.steps} steps (hoist={a.hoist})")
if a.spans:
    for name, (ms, n) in sorted(hp.timer.totals().items(), key=lambda kv: -kv[1][0]):
        print(f"  {name:12s} {ms / a.steps:8.4f} ms/step  ({n // a.steps} spans/step)")
