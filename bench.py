"""bench.py -- LLMRec hot path on B200: train interactions/s (+ full-catalog eval users/s).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our CUDA path, netflix-shaped synthetic
    python bench.py --impl reference --steps 20 --warmup 3    # the reference's CPU algorithm (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" is one full training step of Trainer (sampled batch of
1024 interactions + augmented edges: forward, 8 BPR/prune heads, backward, dense AdamW).
  value      whole-job interactions/s with the batch indices already resident in HBM (CUDA events)
  e2e        the same through Trainer's public API: host sampler -> pinned H2D of the index batch ->
             step -> D2H read of the loss, every step
  roofline   the dominant kernel family of the step: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/llmrec_oracle.py, torch CPU) on a bounded sample of the same workload
  eval       Trainer.test() over every test user (scoring + top-50 + metrics), users/s
Features (704 MB) exceed the 126 MB L2, so consecutive steps cannot be served from cache ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (dataset, n_users, n_items, interactions, dims, embed, weight_size)
    "netflix": ("netflix", 13187, 17366, 68933, (512, 768, 1536), 64, "[64, 64]"),
    "movielens": ("movielens", 12495, 10322, 57960, (512, 768, 1536), 128, "[128,128,128]"),
}


def ensure_dataset(name):
    from llmrec_b200.synth import make_dataset
    ds, nu, ni, ne, dims, _, _ = WORKLOADS[name]
    root = os.path.join(os.environ.get("LLMREC_BENCH_DIR", "/tmp/llmrec_bench"), f"{name}_seed0") + "/"
    marker = os.path.join(root, ".complete")
    if not os.path.exists(marker):
        os.makedirs(root, exist_ok=True)
        make_dataset(root, dataset=ds, n_users=nu, n_items=ni, n_inter=ne, dims=dims, seed=0)
        open(marker, "w").write("ok")
    return root


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], threading.Event(), index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        try:
            self.proc.terminate()
        except Exception:
            pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def step_bytes(tr):
    """Algorithmic bytes per kernel family for one training step (SURVEY.md 8d formulas, fp32, int32 idx)."""
    hp = tr.hot
    nu, ni, d, S, L = hp.nu, hp.ni, hp.d, hp.S, hp.L
    nnz = tr.graph.nnz
    f = hp.feats
    gemms = [(ni, f["image"].shape[1]), (ni, f["text"].shape[1])] + [(ni, v.shape[1]) for v in f["item"].values()] + [(nu, f["user"].shape[1])]
    proj = sum(4 * n * k + 4 * k * d + 4 * n * d for n, k in gemms)

    def spmm(M, N, segs):
        return 4 * nnz + 4 * (M + 1) + 4 * M + segs * (4 * d * N + 4 * d * M)
    fwd = spmm(nu, ni, S + 1) + spmm(ni, nu, S + 2) + spmm(nu, ni, 2) + (spmm(ni, nu, 1) if L >= 2 else 0)
    bwd = spmm(ni, nu, 1) + spmm(nu, ni, S + 2) + spmm(ni, nu, S + 1) + (spmm(nu, ni, 1) + spmm(ni, nu, 1) if L >= 2 else 0)
    n_par = sum(p.numel() for p in tr.hot.opt.params)
    T = 3 + len(hp.keys)
    fuse = 4 * d * (nu + ni) * ((L + 1) + T + 1)
    return {"proj_fwd": proj, "proj_wgrad": proj, "spmm_fwd": fwd, "spmm_bwd": bwd, "adamw": 28 * n_par,
            "fuse_fwd": fuse, "fuse_bwd": fuse + 4 * d * (nu + ni) * T}


def run_ours(a):
    import torch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 or a.workload == "synthetic":
        from llmrec_b200.dist_bench import run_sharded
        return run_sharded(a)
    from llmrec_b200 import main as M, ops
    from llmrec_b200.runtime import set_args
    from llmrec_b200.utility import batch_test
    from llmrec_b200.utility.load_data import Data
    from llmrec_b200.utility.parser import parse_args, resolve_dataset_dir

    ds, nu, ni, ne, dims, embed, wsize = WORKLOADS[a.workload]
    root = ensure_dataset(a.workload)
    args = set_args(parse_args(["--data_path", root, "--dataset", ds, "--debug", "--epoch", "1", "--embed_size", str(embed),
                                "--weight_size", wsize, "--proj_mode", a.proj_mode, "--feat_layout", a.feat_layout, "--host_sampler", a.host_sampler, "--cuda_graph", str(a.graph)]))
    torch.cuda.set_device(0)
    M.set_seed(args.seed)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):          # stdout carries exactly one JSON line
        gen = Data(path=resolve_dataset_dir(args.data_path, args.dataset), batch_size=args.batch_size, sampler=args.host_sampler)
        batch_test.init(gen, args)
        tr = M.Trainer(data_config={}, data_generator=gen)
    hp = tr.hot
    K, W = a.steps, max(a.warmup, 3)

    # ---- device-resident leg ("value") --------------------------------------------------------------
    batches = [tr.sample_batch() for _ in range(W + K)]
    dev_batches = []
    for u, p, n in batches:
        t = torch.tensor([u, p, n], dtype=torch.int32, device="cuda")
        dev_batches.append((t[0], t[1], t[2]))
    step = hp.train_step_graphed if a.graph else hp.train_step
    clocks = ClockSampler(0); clocks.start()      # samples clocks / throttle reasons across the value and e2e legs
    for i in range(W):
        step(*dev_batches[i])
    torch.cuda.synchronize()
    l0 = ops.STATS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        step(*dev_batches[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ops.STATS["launches"] - l0
    n_inter = sum(len(b[0]) for b in batches[W:])
    value = n_inter / (ms / 1e3)
    if a.graph:      # kernels per replayed step = launches of one eager step
        l1 = ops.STATS["launches"]
        snap = hp._snapshot_state(); hp.train_step(*dev_batches[0]); hp._restore_state(snap)
        launches = (ops.STATS["launches"] - l1) * K

    # ---- end-to-end leg through Trainer's API ------------------------------------------------------------
    loss_host = torch.empty(K, dtype=torch.float32).pin_memory()
    for _ in range(3):
        tr.train_next_batch()
    torch.cuda.synchronize()
    e0.record()
    n_e2e = 0
    h2d = 0
    for i in range(K):
        loss, B = tr.train_next_batch()          # the body of Trainer.train()'s loop: host sampler -> pinned staging -> H2D -> step
        loss_host[i:i + 1].copy_(loss, non_blocking=True)
        n_e2e += B; h2d += 3 * 4 * B
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.finish()
    assert bool(torch.isfinite(loss_host).all()), "non-finite loss"

    # ---- per-kernel-family device time (roofline leg): each family's launches of one step, replayed R times from
    #      its own CUDA graph between two events -> pure device time, no host gaps
    u0, p0, n0 = dev_batches[0]
    hp.forward(); hp.loss_and_output_grads(u0, p0, n0); hp.backward()
    torch.cuda.synchronize()
    snap = hp._snapshot_state()
    fams = {"proj_fwd": hp._proj_fwd, "spmm_fwd": hp._prop_fwd, "fuse_fwd": hp._fuse_fwd,
            "loss_heads": lambda: hp.loss_and_output_grads(u0, p0, n0), "fuse_bwd": hp._fuse_bwd, "spmm_bwd": hp._chain_bwd,
            "proj_wgrad": hp._wgrad, "adamw": lambda: hp.opt.step([hp.grads[k] for k in hp._opt_names])}
    R = 10
    fam = {}
    for name, fn in fams.items():
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(R):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        fam[name] = e0.elapsed_time(e1) / R
    hp._restore_state(snap)
    bytes_ = step_bytes(tr)
    hbm, tf, src = peaks()
    top = max((k for k in fam if k in bytes_), key=lambda k: fam[k])
    ach = bytes_[top] / (fam[top] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:                                          # per-launch DRAM bytes of this kernel family from the committed ncu capture
        t = json.load(open(os.path.join(REPO, "profiles", "r1_ncu_traffic.json"))).get(top)
        if t and a.workload == "netflix":
            traffic, traffic_src = int(t["bytes"]), t["source"]
    except Exception:
        pass
    roof = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4),
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": src, "alg_bytes_per_step": bytes_[top], "ms_per_step": round(fam[top], 4),
            "families_ms": {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
            "families_gbs": {k: round(bytes_[k] / (fam[k] * 1e-3) / 1e9, 1) for k in fam if k in bytes_}}

    # ---- eval leg ----------------------------------------------------------------------------------------------
    users = list(gen.test_set.keys())
    tr.test(users, False)                         # warm-up at the timed size (scratch buffers exist afterwards)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = tr.test(users, False)
    torch.cuda.synchronize()
    t_eval = time.perf_counter() - t0
    ev = {"metric": "eval_users_per_sec", "value": round(len(users) / t_eval, 1), "unit": "users/s", "n_users": len(users),
          "n_items": ni, "seconds": round(t_eval, 4), "recall@20": float(res["recall"][1]), "includes": "forward + scoring + top-50 + metrics, host buffers"}

    out = {"metric": "train_interactions_per_sec", "value": round(value, 1), "unit": "interactions/s", "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "impl": "ours",
           "config": {"workload": f"{a.workload}-shaped synthetic {nu}x{ni}, {gen.n_train} train edges, d={embed}, L={len(eval(wsize))}, batch=1024 (+aug edges), feature dims {list(dims)}",
                      "interactions_counted": "sum(len(users)) incl. augmented edges", "l2": "inputs larger than L2 (704 MB of features per step)",
                      "proj_mode": a.proj_mode, "feat_layout": a.feat_layout, "host_sampler": a.host_sampler, "cuda_graph": bool(a.graph)},
           "e2e": {"value": round(n_e2e / (ms_e2e / 1e3), 1), "unit": "interactions/s", "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": 4,
                   "ms_per_step": round(ms_e2e / K, 4)},
           "gpu_launches": launches, "clocks": clk, "roofline": roof, "eval": ev}
    if not a.no_cpu:
        out["cpu_baseline"] = cpu_baseline(a, steps=a.cpu_steps, eval_users=a.cpu_eval_users)
    return out


def cpu_baseline(a, steps, eval_users, warmup=1):
    """The oracle port timed on the host cores: `steps` training steps + eval of `eval_users` users."""
    import torch
    from oracle import llmrec_oracle as O
    ds, nu, ni, ne, dims, embed, wsize = WORKLOADS[a.workload]
    root = ensure_dataset(a.workload)
    from llmrec_b200.synth import DATASET_DIR
    data = O.load_dataset(os.path.join(root, DATASET_DIR[ds]))
    cores = os.cpu_count()
    cfg = O.OracleConfig(embed_size=embed, weight_size=tuple(eval(wsize)))
    O.set_seed(cfg.seed)
    tr = O.OracleTrainer(data, cfg)
    # torch's intra-op pool oversubscribes badly on many-core hosts (128 threads: 5 s/step vs 0.25 s at 8): use the
    # thread count that runs one step fastest ("all the host threads it can use" without thrashing)
    best_t, best_dt = None, None
    for nt in sorted({min(cores, c) for c in (8, 16, 32, 64)}):
        torch.set_num_threads(nt)
        b = O.sample_batch(data, cfg)
        tr.step(*b)
        t0 = time.perf_counter(); tr.step(*b); dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best_t, best_dt = nt, dt
    torch.set_num_threads(best_t)
    for _ in range(warmup):
        tr.step(*O.sample_batch(data, cfg))
    t0 = time.perf_counter(); n = 0
    for _ in range(steps):
        u, p, ng = O.sample_batch(data, cfg)
        tr.step(u, p, ng)
        n += len(u)
    dt = time.perf_counter() - t0
    users = list(data.test_set.keys())[:eval_users]
    t1 = time.perf_counter()
    if users:
        tr.test(users, faithful=True)
    de = time.perf_counter() - t1
    return {"value": round(n / dt, 1), "unit": "interactions/s", "cores": best_t, "host_cores": cores, "kind": "port",
            "sample": f"{steps} training steps ({dt:.1f} s) of the same workload, torch {torch.__version__} CPU with {torch.get_num_threads()} threads",
            "ms_per_step": round(dt / steps * 1e3, 2),
            "eval": {"value": round(len(users) / de, 1) if users else None, "unit": "users/s", "sample": f"{len(users)} test users, per-user heapq ranking ({de:.1f} s)"}}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    ds, nu, ni, ne, dims, embed, wsize = WORKLOADS[a.workload]
    cb = cpu_baseline(a, steps=max(1, a.steps), eval_users=a.cpu_eval_users, warmup=max(1, min(a.warmup, 3)))
    return {"metric": "train_interactions_per_sec", "value": cb["value"], "unit": "interactions/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{a.workload}-shaped synthetic {nu}x{ni}, d={embed}, batch=1024 (+aug edges); reference algorithm as the CPU oracle port (the Python reference cannot travel to the GPU box)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "eval": cb["eval"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix", choices=list(WORKLOADS) + ["synthetic"])
    ap.add_argument("--pieces", type=int, default=1, help="N>1: item-row pieces of the exchange SpMMs (all-reduce of piece k overlaps SpMM of piece k+1)")
    ap.add_argument("--item-sharded", dest="item_sharded", type=int, default=0, help="N>1: reduce-scatter / row-local work / all-gather form of the item-side exchanges, AdamW of the item table sharded by item")
    ap.add_argument("--n1-base", dest="n1_base", type=int, default=1, help="N>1: also time the same synthetic workload on rank 0 alone")
    ap.add_argument("--eval-users", dest="eval_users", type=int, default=102400, help="users ranked in the synthetic eval leg")
    ap.add_argument("--syn-scale", dest="syn_scale", type=float, default=1.0, help="size factor of the 10M x 1M x 200M synthetic graph")
    ap.add_argument("--proj_mode", default="3xtf32")
    ap.add_argument("--feat_layout", default="rows", choices=["rows", "panels"])
    ap.add_argument("--host_sampler", default="native")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--cpu-steps", dest="cpu_steps", type=int, default=24)
    ap.add_argument("--cpu-eval-users", dest="cpu_eval_users", type=int, default=1500)
    a = ap.parse_args()
    out = run_reference(a) if a.impl == "reference" else run_ours(a)
    if out is not None and int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
