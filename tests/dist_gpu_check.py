"""Sharded ID-only path vs the single-GPU engine on the same graph, same init, same batches.
    python tests/dist_gpu_check.py                                  (world 1)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Prints DIST_CHECK_OK on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from llmrec_b200.dist import ShardedGraph, ShardedHotPath, shard_bounds
from llmrec_b200.engine import HotPath, HotPathConfig
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
    torch.manual_seed(0)
    Eu, Ei = torch.randn(nu, d) * 0.1, torch.randn(ni, d) * 0.1
    cfg = HotPathConfig(embed_size=d, n_layers=L, batch_size=256)
    b = shard_bounds(nu, world)
    lo, hi = b[rank], b[rank + 1]
    mine = e[(e[:, 0] >= lo) & (e[:, 0] < hi)]
    g = ShardedGraph(torch.from_numpy(mine[:, 0] - lo).to(dev), torch.from_numpy(mine[:, 1]).to(dev), hi - lo, ni, tile_nnz=32, pieces=(2 if world > 1 else 1))
    sh = ShardedHotPath(g, Eu[lo:hi].clone().to(dev), Ei.clone().to(dev), cfg, lo, item_sharded=os.environ.get("LLMREC_DIST_ITEM_SHARDED") == "1",
                        demand=os.environ.get("LLMREC_DIST_DEMAND") == "1")
    # reference: single-GPU engine on the full graph (every rank builds it; small)
    R = sp.csr_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(nu, ni))
    bg = BipartiteGraph(R, dev, tile_nnz=32)
    params = {"user_id_embedding.weight": Eu.clone().to(dev), "item_id_embedding.weight": Ei.clone().to(dev)}
    hp = HotPath((bg.ui, bg.iu, bg.uiT, bg.iuT), params, None, cfg)
    hp.set_optimizer(lr=1e-3); sh.set_lr(1e-3)
    ok = True
    for step in range(3):
        users = torch.from_numpy(rng.integers(0, nu, 280).astype(np.int32)).to(dev)
        pos = torch.from_numpy(rng.integers(0, ni, 280).astype(np.int32)).to(dev)
        neg = torch.from_numpy(rng.integers(0, ni, 280).astype(np.int32)).to(dev)
        l1 = float(hp.train_step(users, pos, neg)); l2 = float(sh.train_step(users, pos, neg))
        ok &= abs(l1 - l2) < 1e-5 * max(1.0, abs(l1))
        rows = torch.cat([pos, neg]).long()           # the training step fuses I on the batch rows only
        if sh.demand:                                 # ... and, in demand mode, U on the batch users only (compact blocks)
            ok &= bool(torch.allclose(sh.Ub, hp.U[users.long()], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(sh.Ib, hp.I[rows], rtol=1e-4, atol=1e-6))
        else:
            ok &= bool(torch.allclose(sh.U, hp.U[lo:hi], rtol=1e-4, atol=1e-6))
            ok &= bool(torch.allclose(sh.I[rows], hp.I[rows], rtol=1e-4, atol=1e-6))
    hp.forward(); sh.forward()                        # full forward (eval form)
    ok &= bool(torch.allclose(sh.U, hp.U[lo:hi], rtol=1e-4, atol=1e-6)) and bool(torch.allclose(sh.I, hp.I, rtol=1e-4, atol=1e-6))
    ok &= bool(torch.allclose(sh.E_u, params["user_id_embedding.weight"][lo:hi], rtol=1e-4, atol=1e-6))
    ok &= bool(torch.allclose(sh.E_i, params["item_id_embedding.weight"], rtol=1e-4, atol=1e-6))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK_OK" if float(flag) == 1.0 else "DIST_CHECK_FAILED", "world", world, "loss", l1, l2, flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
