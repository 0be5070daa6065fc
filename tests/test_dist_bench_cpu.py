"""bench.py's N > 1 leg end to end on CPU: llmrec_b200.dist_bench.run_sharded at world size 1 and 2 under gloo with torch stand-ins
for the kernels (tests/ops_emulator.py), INCLUDING the `--n1-base 1` branch that frees the sharded engine and re-times the
workload on rank 0 alone -- the branch whose UnboundLocalError took down every multi-GPU bench line of round 1.  The output must
be the dictionary the driver parses: metric / value / ms_per_step / e2e / roofline / eval / same_workload_1gpu, JSON-serialisable."""
import argparse
import json
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _args(**kw):
    a = argparse.Namespace(steps=2, warmup=3, pieces=1, item_sharded=0, n1_base=1, eval_users=64, syn_scale=2e-4, min_seconds=0.0,
                           max_blocks=2, demand=0)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _worker(rank, world, port, kw, out):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import ops_emulator
    ops_emulator.install()
    from llmrec_b200.dist_bench import run_sharded
    res = run_sharded(_args(**kw), dev=torch.device("cpu"))
    out[rank] = json.dumps(res)                  # must serialise (rank 0: the line; others: null)
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kw", [(1, {}), (2, {}), (2, {"item_sharded": 1, "pieces": 1}), (2, {"pieces": 2, "n1_base": 0}), (2, {"demand": 1})])
def test_run_sharded_builds_the_bench_line(world, kw):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kw, out), nprocs=world, join=True)
    res = {r: json.loads(v) for r, v in out.items()}
    assert all(res[r] is None for r in range(1, world))
    line = res[0]
    assert line["metric"] == "train_interactions_per_sec" and line["n_gpus"] == world and line["scaling"] == "strong"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 3 * 4 * 1126
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["achieved"] > 0
    assert line["eval"]["value"] > 0 and line["eval"]["all_ranked"] and line["eval"]["shots"] >= 5
    assert line["config"]["item_sharded"] == bool(kw.get("item_sharded", 0))
    if world > 1 and kw.get("n1_base", 1):
        base = line["same_workload_1gpu"]
        assert base["n_gpus"] == 1 and base["value"] > 0 and line["scaling_base"] == base and line["speedup_vs_1gpu"] > 0
    else:
        assert line["same_workload_1gpu"] is None
