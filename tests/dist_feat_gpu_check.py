"""Sharded FULL path (dist_feat.ShardedFeatureHotPath) vs the single-GPU engine on the same graph, features, init and batches.
    python tests/dist_feat_gpu_check.py                                  (world 1)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/dist_feat_gpu_check.py
Prints DIST_FEAT_CHECK_OK on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from llmrec_b200.dist import ShardedGraph, shard_bounds
from llmrec_b200.dist_feat import ShardedFeatureHotPath
from llmrec_b200.engine import HotPath, HotPathConfig, PARAM_ORDER
from llmrec_b200.graph import BipartiteGraph


def main():
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    nu, ni, d, L = 3001, 1700, 64, 2
    w = 1.0 / (np.arange(ni) + 4.0); w /= w.sum()
    e = np.unique(np.stack([np.concatenate([np.arange(nu), rng.integers(0, nu, 30000)]), rng.choice(ni, size=nu + 30000, p=w)], 1), axis=0)
    gen = torch.Generator().manual_seed(0)
    keys = ["a", "b", "c"]
    dims = dict(image=64, text=96, user=160, item=160)
    full = {"user_id_embedding.weight": torch.randn(nu, d, generator=gen) * 0.1, "item_id_embedding.weight": torch.randn(ni, d, generator=gen) * 0.1}
    for name in ("image", "text", "user", "item"):
        full[name + "_trans.weight"] = torch.randn(d, dims[name], generator=gen) / dims[name] ** 0.5
        full[name + "_trans.bias"] = torch.randn(d, generator=gen) * 0.1
    feats = dict(image=torch.randn(ni, dims["image"], generator=gen), text=torch.randn(ni, dims["text"], generator=gen),
                 user=torch.randn(nu, dims["user"], generator=gen), item={k: torch.randn(ni, dims["item"], generator=gen) for k in keys})
    cfg = HotPathConfig(embed_size=d, n_layers=L, batch_size=256)
    ub, ib = shard_bounds(nu, world), shard_bounds(ni, world)
    lo, hi, ilo, ihi = ub[rank], ub[rank + 1], ib[rank], ib[rank + 1]
    mine = e[(e[:, 0] >= lo) & (e[:, 0] < hi)]
    g = ShardedGraph(torch.from_numpy(mine[:, 0] - lo).to(dev), torch.from_numpy(mine[:, 1]).to(dev), hi - lo, ni, tile_nnz=32)
    p_sh = {k: (v[lo:hi] if k == "user_id_embedding.weight" else v).clone().to(dev) for k, v in full.items()}
    f_sh = dict(image=feats["image"][ilo:ihi].to(dev), text=feats["text"][ilo:ihi].to(dev), user=feats["user"][lo:hi].to(dev),
                item={k: v[ilo:ihi].to(dev) for k, v in feats["item"].items()})
    sh = ShardedFeatureHotPath(g, p_sh, f_sh, cfg, lo, ilo)
    R = sp.csr_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(nu, ni))
    bg = BipartiteGraph(R, dev, tile_nnz=32)
    p_one = {k: v.clone().to(dev) for k, v in full.items()}
    f_one = dict(image=feats["image"].to(dev), text=feats["text"].to(dev), user=feats["user"].to(dev), item={k: v.to(dev) for k, v in feats["item"].items()})
    hp = HotPath((bg.ui, bg.iu, bg.uiT, bg.iuT), p_one, f_one, cfg)
    hp.set_optimizer(lr=1e-3); sh.set_lr(1e-3)
    ok = True
    tol = dict(rtol=2e-4, atol=2e-6)
    for step in range(3):
        users = torch.from_numpy(rng.integers(0, nu, 280).astype(np.int32)).to(dev)
        pos = torch.from_numpy(rng.integers(0, ni, 280).astype(np.int32)).to(dev)
        neg = torch.from_numpy(rng.integers(0, ni, 280).astype(np.int32)).to(dev)
        l1 = float(hp.train_step(users, pos, neg)); l2 = float(sh.train_step(users, pos, neg))
        ok &= abs(l1 - l2) < 2e-5 * max(1.0, abs(l1))
    for k in PARAM_ORDER:
        ok &= bool(torch.allclose(p_sh[k], p_one[k][lo:hi] if k == "user_id_embedding.weight" else p_one[k], **tol))
    hp.forward(); sh.forward()
    ok &= bool(torch.allclose(sh.U, hp.U[lo:hi], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(sh.I, hp.I, rtol=1e-4, atol=1e-6))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_FEAT_CHECK_OK" if float(flag) == 1.0 else "DIST_FEAT_CHECK_FAILED", "world", world, "loss", l1, l2, flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
